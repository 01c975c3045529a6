"""The UNMODIFIED reference libbfm, compiled by oracle/Makefile into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

It shares our C ABI, so it is driven through the same ctypes declarations and the same object
model as the product (bfm_b200.api) - only the .so differs.
"""

from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbfm_ref.so")
REFERENCE_TREE = os.environ.get("BFM_REFERENCE", "/root/reference")

_binding = None


def build():
	"""(re)build from the reference tree when it is present; otherwise keep the prebuilt file"""

	subprocess.run(["make", "-C", _HERE, "ref", f"REFERENCE={REFERENCE_TREE}"], check=True, capture_output=True)


def available() -> bool:
	if not os.path.exists(LIB_PATH) and os.path.isdir(REFERENCE_TREE):
		build()

	return os.path.exists(LIB_PATH)


def binding():
	global _binding

	if _binding is None:
		if not available():
			raise RuntimeError("oracle/_ref/libbfm_ref.so is missing and the reference tree is not present")

		from bfm_b200.api import Binding

		_binding = Binding(LIB_PATH)

	return _binding
