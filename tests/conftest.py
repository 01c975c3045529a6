import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

for p in (ROOT, os.path.join(ROOT, "tests")):
	if p not in sys.path:
		sys.path.insert(0, p)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lib():
	"""our libbfm.so; built in-tree on demand (nvcc cross-compiles without a GPU)"""

	from bfm_b200 import api, build

	if not os.path.exists(api.LIB_PATH):
		build.build()

	return api.default_binding()


@pytest.fixture(scope="session")
def ref():
	"""the unmodified reference compiled by oracle/Makefile; absent only if neither the prebuilt
	oracle/_ref/libbfm_ref.so nor the reference tree is there"""

	from oracle import ref as ref_mod

	if not ref_mod.available():
		pytest.skip("oracle/_ref/libbfm_ref.so not available")

	return ref_mod.binding()


@pytest.fixture(scope="session")
def golden():
	import cases

	return cases.golden()
