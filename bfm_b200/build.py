"""In-tree build of libbfm.so (host C + sm_100a CUDA) with gcc/nvcc - no CMake, no JIT cache.

    python -m bfm_b200.build [--force] [--verbose]

Objects go to bfm_b200/csrc/_obj/, the library to bfm_b200/lib/libbfm.so (SONAME libbfm.so.1, the
reference's, libbfm/CMakeLists.txt:23-24).  nvcc cross-compiles for sm_100a without a GPU present.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(SRC, "_obj")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libbfm.so")

C_SOURCES = ["core.c", "mesh.c", "ez.c", "matrix.c", "perm.c", "plan.c", "renumber.c", "partition.c", "coarse.c", "hier.c", "csr.c", "job.c"]

# assembly.cu must not contract a*b+c into an FMA: bit-for-bit parity with the reference's gcc/x86-64
# arithmetic (see the header of assembly.cu).  The solver is free to use FMAs.
CU_SOURCES = {"context.cu": ["-Xcompiler", "-fopenmp"], "assembly.cu": ["-fmad=false"], "solver.cu": [], "dist.cu": [], "batch.cu": [], "symbolic.cu": []}

INCLUDES = ["-I" + os.path.join(ROOT, "include"), "-I" + SRC]

CFLAGS = ["-O2", "-g", "-fPIC", "-std=gnu11", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Wno-sign-compare"]

NVCC_FLAGS = [
	"-gencode", "arch=compute_100a,code=sm_100a",
	"-lineinfo", "-O3", "-std=c++17",
	"-Xcompiler", "-fPIC",
	"-Xptxas", "-v",
]


def _nvcc() -> str:
	for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
		if cand and os.path.exists(cand):
			return cand

	raise RuntimeError("nvcc not found")


def _run(cmd: list[str], verbose: bool, log: list[str]):
	if verbose:
		print(" ".join(cmd), flush=True)

	proc = subprocess.run(cmd, capture_output=True, text=True)
	log.append("$ " + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)

	if proc.returncode != 0:
		sys.stderr.write(proc.stdout + proc.stderr)
		raise RuntimeError("build step failed: " + " ".join(cmd))

	if verbose:
		sys.stdout.write(proc.stdout + proc.stderr)


def _stale(target: str, deps: list[str]) -> bool:
	if not os.path.exists(target):
		return True

	t = os.path.getmtime(target)
	return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
	os.makedirs(OBJ, exist_ok=True)
	os.makedirs(LIB_DIR, exist_ok=True)

	headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
	headers += [os.path.join(ROOT, "include", "bfm_b200.h"), os.path.join(ROOT, "include", "bfm", "libbfm.h"), os.path.abspath(__file__)]

	cc = os.environ.get("CC") or shutil.which("gcc") or "gcc"
	nvcc = _nvcc()
	log: list[str] = []
	objects = []

	for name in C_SOURCES:
		src = os.path.join(SRC, name)
		obj = os.path.join(OBJ, name + ".o")
		objects.append(obj)

		if force or _stale(obj, [src] + headers):
			_run([cc, *CFLAGS, *INCLUDES, "-c", src, "-o", obj], verbose, log)

	for name, extra in CU_SOURCES.items():
		src = os.path.join(SRC, name)

		if not os.path.exists(src):
			continue

		obj = os.path.join(OBJ, name + ".o")
		objects.append(obj)

		if force or _stale(obj, [src] + headers):
			_run([nvcc, *NVCC_FLAGS, *extra, *INCLUDES, "-c", src, "-o", obj], verbose, log)

	if force or _stale(LIB, objects):
		_run([
			nvcc, "-shared", "-o", LIB, *objects,
			"-gencode", "arch=compute_100a,code=sm_100a",
			"-Xcompiler", "-fopenmp", "-Xlinker", "-soname=libbfm.so.1", "-Xlinker", "-Bsymbolic", "-lm", "-ldl",
		], verbose, log)

	with open(os.path.join(OBJ, "build.log"), "w") as f:
		f.write("\n".join(log))

	return LIB


if __name__ == "__main__":
	print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
