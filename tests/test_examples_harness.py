"""SURVEY.md 8(f).1 - the reference's own scripts, UNMODIFIED, on this library (tools/examples_harness).

Needs the reference checkout and cffi: /root/reference in the build container; on the GPU box the copy of its scripts,
pybfm sources, headers and meshes that __graft_entry__.build() stages under the git-ignored baseline/_ref/.  The harness runs the
reference's binding generator against our headers and library, then the script with a pyglet stand-in:
  - on the compiled reference library (oracle/_ref) the scripts run to completion on the CPU and
    lepl1110.py rewrites the reference's golden U.txt / V.txt byte for byte - the harness adds nothing;
  - on OUR library, without a CUDA device, they get exactly as far as `lib.bfm_sim_run` and stop there
    with libbfm's loud "no usable CUDA device" (there is no CPU path); with a device they complete and
    the files agree with the golden ones to the printed digits.
"""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tools", "examples_harness", "run_reference_example.py")
sys.path.insert(0, os.path.dirname(HARNESS))

import run_reference_example  # noqa: E402

REFERENCE = run_reference_example.default_reference()

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "pybfm")), reason="reference checkout not present")


def _run(workdir, script, *args, library=None):
	pytest.importorskip("cffi")

	cmd = [sys.executable, HARNESS, "--reference", REFERENCE, "--workdir", str(workdir)]

	if library:
		cmd += ["--library", library]

	env = dict(os.environ)
	env.pop("BFM_QUIET", None)

	return subprocess.run(cmd + [script, *args], capture_output=True, text=True, timeout=600, env=env)


def _have_gpu(lib) -> bool:
	from bfm_b200 import ext

	return ext.device_available(lib)


def test_harness_is_transparent_on_the_reference_library(tmp_path, ref):
	from oracle import ref as ref_mod

	proc = _run(tmp_path, "lepl1110.py", "meshes/8.lepl1110", "problems/problem.txt", library=ref_mod.LIB_PATH)
	assert proc.returncode == 0, proc.stderr[-3000:]

	for name in ("U.txt", "V.txt"):
		assert (tmp_path / "data" / name).read_bytes() == open(os.path.join(REFERENCE, "data", name), "rb").read()


@pytest.mark.parametrize("script,args", [
	("lepl1110.py", ("meshes/8.lepl1110", "problems/problem.txt")),
	("examples/benchmark.py", ()),
	("examples/deformation.py", ()),
])
def test_reference_scripts_stop_loudly_without_a_device(script, args, tmp_path, lib):
	if _have_gpu(lib):
		pytest.skip("a device is present: see test_reference_scripts_run_unmodified_on_the_gpu")

	proc = _run(tmp_path, script, *args)

	# no device here: everything up to the hot path worked through the reference's own cffi binding
	# (mesh readers, problem parser, object model, GL-free instance set-up), and the hot path said why it stopped
	assert proc.returncode != 0
	assert "lib.bfm_sim_run" in proc.stderr and "AssertionError" in proc.stderr
	assert "no usable CUDA device" in proc.stderr and "no CPU fallback" in proc.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("script,args", [
	("lepl1110.py", ("meshes/8.lepl1110", "problems/problem.txt")),
	("examples/benchmark.py", ()),
	("examples/deformation.py", ()),
])
def test_reference_scripts_run_unmodified_on_the_gpu(script, args, tmp_path, lib):
	"""BASELINE.json north_star: "lepl1110.py, examples/deformation.py and examples/benchmark.py run unchanged" - through
	the reference's own cffi binding (pybfm/bfm/gen_libbfm.py), on the B200; lepl1110.py must write the reference's golden
	U.txt / V.txt to the printed 8 digits (pybfm/bfm/sim.py:33-34 -> bfm_sim_run -> ez.c:211-229)"""

	assert _have_gpu(lib), lib.lib.bfmx_device_error()

	proc = _run(tmp_path, script, *args)

	assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-3000:]

	if script == "lepl1110.py":
		import numpy as np

		def numbers(text):  # "%14.7e" x 3 per line, no separator guaranteed (ez.c:218-225)
			body = text.split("\n", 1)[1].replace("\n", "")
			return [float(body[i:i + 14]) for i in range(0, len(body), 14)]

		for name in ("U.txt", "V.txt"):
			got = numbers((tmp_path / "data" / name).read_text())
			want = numbers(open(os.path.join(REFERENCE, "data", name)).read())
			assert len(got) == len(want) == 335 and np.allclose(got, want, rtol=2e-7, atol=1e-16)

	if script == "examples/benchmark.py":
		assert "Average time taken to simulate" in proc.stdout  # examples/benchmark.py:18

	if script == "examples/deformation.py":
		assert (tmp_path / "index.html").exists()  # Bfm.export wrote the WebGL page with the displacements inlined
