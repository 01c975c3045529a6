"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): the row-partitioned path, one process per
GPU under torchrun with NCCL over NVLink, against the reference outputs - see tests/dist_gpu_worker.py."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

CASES = ["lepl8", "lepl8_all_kinds", "bridge", "plate_80x20", "plate_q4_24x6", "plate_neumann_16x4", "plate_funky_20x5", "lepl8_axisym", "gear60", "plate_jitter_40x10"]


def _gpu_count() -> int:
	try:
		out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
		return sum(1 for line in out.splitlines() if line.startswith("GPU "))
	except (OSError, subprocess.TimeoutExpired):
		return 0


@pytest.mark.parametrize("world,exchange", [(2, "peer"), (2, "nccl"), (2, "peer-renumbered"), (4, "peer"), (8, "peer")])
def test_partitioned_run_matches_reference(world, exchange, tmp_path):
	"""exchange: the per-iteration halo / dot-product / coarse exchanges over NVLink peer memory (CUDA IPC
	mailboxes, posted from inside the kernels - the default) or over NCCL (BFM_P2P=0)"""

	if _gpu_count() < world:
		pytest.skip(f"needs {world} GPUs")

	out = tmp_path / "report.json"
	cmd = [
		sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
		"--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
		os.path.join(ROOT, "tests", "dist_gpu_worker.py"), ",".join(CASES), str(out),
	]

	env = dict(os.environ, BFM_P2P="0" if exchange == "nccl" else "1")

	if exchange == "peer-renumbered": # every mesh partitioned on its internal Morton numbering (bfm_b200/csrc/renumber.c)
		env["BFM_RENUMBER"] = "1"

	proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
	assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-6000:]

	reports = json.loads(out.read_text())
	assert len(reports) == world

	for rank, report in enumerate(reports):
		for name in CASES:
			r = report[name]
			assert r["converged"] == 1 and r["rel_residual"] <= 1e-12, (rank, name, r)
			assert r["err"] <= 1e-9, (rank, name, r)
			assert r["same_on_all_ranks"] and r["b_bitwise"] and r["staged_equals_run"], (rank, name, r)
			assert r["n_ranks"] == world and r["peer_memory"] == (0 if exchange == "nccl" else 1)
