#pragma once
// forwarder: the declarations the reference keeps in bfm/rule.h live in bfm/libbfm.h
#include <bfm/libbfm.h>
