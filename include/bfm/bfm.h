#pragma once
// forwarder: the declarations the reference keeps in bfm/bfm.h live in bfm/libbfm.h
#include <bfm/libbfm.h>
