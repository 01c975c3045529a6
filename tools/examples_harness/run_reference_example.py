#!/usr/bin/env python
"""Run one of the REFERENCE's own scripts, unmodified, on this repository's libbfm.so.

    python tools/examples_harness/run_reference_example.py [--reference /root/reference] [--workdir DIR] \
        lepl1110.py meshes/8.lepl1110 problems/problem.txt
    python tools/examples_harness/run_reference_example.py examples/benchmark.py
    python tools/examples_harness/run_reference_example.py examples/deformation.py

What it does - nothing of the reference is copied into this repository, the run directory is scratch:

  1. builds a run directory that mirrors the reference checkout with SYMLINKS (scripts, pybfm sources,
     meshes, problems, shaders, web), real `data/` and `meshes/` directories for the files the scripts
     write, and a generated stand-in for the missing `meshes/terrain.obj` (.MISSING_LARGE_BLOBS:1);
  2. runs the reference's OWN binding generator, pybfm/bfm/gen_libbfm.py, there: it cdef's the reference's
     headers and compiles the cffi module against OUR headers (include/bfm/*.h via C_INCLUDE_PATH) and links
     OUR library (-lbfm via LIBRARY_PATH) - the drop-in boundary of INTEGRATION.md, exercised for real;
  3. runs the script with a pyglet stand-in on PYTHONPATH (no display on a compute box) and our libbfm.so.1
     on LD_LIBRARY_PATH.

Exit status is the script's.  Without a CUDA device the script stops at `assert not lib.bfm_sim_run(...)`
(pybfm/bfm/sim.py:34) with libbfm's "no usable CUDA device" message on stderr - there is no CPU path.
"""

from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


STAGED = os.path.join(ROOT, "baseline", "_ref", "reference")


def default_reference() -> str:
	"""the reference checkout: $BFM_REFERENCE, /root/reference (the build container), else the copy __graft_entry__.build()
	staged under the git-ignored baseline/_ref/ so that it travels to the GPU box"""

	for cand in (os.environ.get("BFM_REFERENCE"), "/root/reference", STAGED):
		if cand and os.path.isdir(os.path.join(cand, "pybfm")):
			return cand

	return "/root/reference"


def stage_reference(reference: str = "/root/reference", target: str = STAGED) -> str | None:
	"""copy what the reference's scripts need at run time (scripts, pybfm sources, headers, meshes, problems, golden
	U/V, shaders, web: 2 MB, no library sources) into baseline/_ref/reference.  The directory is git-ignored - nothing
	of the reference enters this repository's history - but not gpurun-ignored, so the examples can run on the GPU box,
	where /root/reference does not exist."""

	if not os.path.isdir(os.path.join(reference, "pybfm")):
		return None

	for name in ("examples", "problems", "shaders", "web", "data", "meshes", os.path.join("pybfm", "bfm"), os.path.join("libbfm", "src", "bfm")):
		shutil.copytree(os.path.join(reference, name), os.path.join(target, name), dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.so", "*.c", "*.o"))

	shutil.copyfile(os.path.join(reference, "lepl1110.py"), os.path.join(target, "lepl1110.py"))

	for root, dirs, files in os.walk(target):
		for name in dirs + files:
			os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)

	return target


def _link(src: str, dst: str):
	os.makedirs(os.path.dirname(dst), exist_ok=True)

	if not os.path.lexists(dst):
		os.symlink(src, dst)


def terrain_stand_in(path: str):
	"""a small 3-D triangulated strip under the bridge: nodes with |z| <= 0.04 give the cross-section that
	examples/deformation.py:38-56 samples for its boundary conditions"""

	nx, nz = 60, 3
	lines = []

	for k in range(nz):
		for i in range(nx):
			x = -3.0 + 6.0 * i / (nx - 1)
			y = -0.45 + 0.12 * x ** 2  # a valley whose flanks rise above the feet of bridge-dam.obj at both ends (27 clamped nodes):
			                           # with no node below the terrain the script would set up a singular, unconstrained system
			z = -0.03 + 0.03 * k
			lines.append(f"v {x:.6f} {y:.6f} {z:.6f}")

	for k in range(nz - 1):
		for i in range(nx - 1):
			a = k * nx + i + 1
			lines.append(f"f {a} {a + 1} {a + nx + 1}")
			lines.append(f"f {a} {a + nx + 1} {a + nx}")

	with open(path, "w") as f:
		f.write("\n".join(lines) + "\n")


def build_rundir(reference: str, workdir: str, library: str | None = None) -> str:
	os.makedirs(workdir, exist_ok=True)

	# read-only parts: plain symlinks
	for name in ("lepl1110.py", "examples", "problems", "shaders", "web"):
		_link(os.path.join(reference, name), os.path.join(workdir, name))

	# the reference's headers are the cdef text gen_libbfm.py reads (libbfm/src/bfm/*.h)
	_link(os.path.join(reference, "libbfm", "src", "bfm"), os.path.join(workdir, "libbfm", "src", "bfm"))

	# pybfm: real directories (the compiled cffi module lands in pybfm/bfm/), sources symlinked one by one
	for name in os.listdir(os.path.join(reference, "pybfm", "bfm")):
		if name.endswith(".py"):
			_link(os.path.join(reference, "pybfm", "bfm", name), os.path.join(workdir, "pybfm", "bfm", name))

	# scripts import `bfm`; the generated module is named pybfm.bfm.libbfm and imported as bfm.libbfm
	# directories the scripts write into
	os.makedirs(os.path.join(workdir, "data"), exist_ok=True)

	for name in os.listdir(os.path.join(reference, "meshes")):
		if name != "cross.py":  # a cache examples/deformation.py rebuilds from terrain.obj
			_link(os.path.join(reference, "meshes", name), os.path.join(workdir, "meshes", name))

	terrain = os.path.join(workdir, "meshes", "terrain.obj")

	if not os.path.exists(terrain):
		terrain_stand_in(terrain)

	# our library under the name the linker (-lbfm) and the loader (SONAME libbfm.so.1) look for
	lib = os.path.abspath(library) if library else os.path.join(ROOT, "bfm_b200", "lib", "libbfm.so")

	if not os.path.exists(lib):
		raise SystemExit(f"{lib} is missing: python -c 'import __graft_entry__ as g; g.build()'")

	_link(lib, os.path.join(workdir, "lib", "libbfm.so"))
	_link(lib, os.path.join(workdir, "lib", "libbfm.so.1"))

	return workdir


def generate_binding(workdir: str, verbose: bool):
	"""pybfm/bfm/gen_libbfm.py, unmodified, from the run directory"""

	target_dir = os.path.join(workdir, "pybfm", "bfm")

	if any(name.startswith("libbfm.") and name.endswith(".so") for name in os.listdir(target_dir)):
		return

	env = dict(os.environ)
	env["C_INCLUDE_PATH"] = os.path.join(ROOT, "include") + os.pathsep + env.get("C_INCLUDE_PATH", "")
	env["LIBRARY_PATH"] = os.path.join(workdir, "lib") + os.pathsep + env.get("LIBRARY_PATH", "")

	proc = subprocess.run([sys.executable, os.path.join("pybfm", "bfm", "gen_libbfm.py")], cwd=workdir, env=env, capture_output=True, text=True)

	if verbose or proc.returncode != 0:
		sys.stderr.write(proc.stdout[-4000:] + proc.stderr[-4000:])

	if proc.returncode != 0:
		raise SystemExit("the reference's binding generator failed against our headers / library")

	# cffi writes the module next to where set_source's dotted name points: pybfm/bfm/libbfm*.so
	built = [name for name in os.listdir(target_dir) if name.startswith("libbfm.") and name.endswith(".so")]

	if not built:
		raise SystemExit("gen_libbfm.py produced no extension module in pybfm/bfm/")


def run_script(workdir: str, script: str, args: list[str]) -> int:
	env = dict(os.environ)
	env["PYTHONPATH"] = os.pathsep.join([os.path.join(workdir, "pybfm"), os.path.join(HERE, "stubs"), env.get("PYTHONPATH", "")])
	env["LD_LIBRARY_PATH"] = os.path.join(workdir, "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")

	return subprocess.run([sys.executable, script, *args], cwd=workdir, env=env).returncode


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--reference", default=default_reference())
	ap.add_argument("--workdir", default=None, help="run directory (default: a fresh temporary directory)")
	ap.add_argument("--library", default=None, help="libbfm to run on (default: ours; oracle/_ref/libbfm_ref.so checks the harness itself on the CPU)")
	ap.add_argument("--keep", action="store_true")
	ap.add_argument("--verbose", action="store_true")
	ap.add_argument("script")
	ap.add_argument("args", nargs="*")
	ns = ap.parse_args()

	if not os.path.isdir(os.path.join(ns.reference, "pybfm")):
		raise SystemExit(f"reference checkout not found at {ns.reference}")

	workdir = ns.workdir or tempfile.mkdtemp(prefix="bfm_rundir_")

	try:
		build_rundir(ns.reference, workdir, ns.library)
		generate_binding(workdir, ns.verbose)
		rc = run_script(workdir, ns.script, ns.args)
	finally:
		if ns.workdir is None and not ns.keep:
			shutil.rmtree(workdir, ignore_errors=True)

	sys.exit(rc)


if __name__ == "__main__":
	main()
