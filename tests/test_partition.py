"""The multi-GPU path's HOST side (bfm_b200/csrc/partition.c), on CPU: structure of the row partition,
agreement of the two ends of every halo link, and a numpy emulation of the partitioned PCG against the
oracle - serially for several world sizes, and with one process per rank over gloo (world_size 2)."""

import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import dist_emulation as emu
from bfm_b200 import api, ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MESHES = ["plate_40x10", "plate_q4_24x6", "bridge", "lepl8"]


def _check_partition(mesh, world):
	nn = mesh.n_nodes
	elems = mesh.elems_array.astype(np.int64)
	parts = [ext.partition(mesh, r, world) for r in range(world)]

	# owned ranges tile the node set
	assert parts[0]["first_node"] == 0 and parts[-1]["end_node"] == nn
	assert all(parts[r]["end_node"] == parts[r + 1]["first_node"] for r in range(world - 1))

	for r, p in enumerate(parts):
		l2g = p["l2g"].astype(np.int64)
		lo, hi = int(p["first_node"]), int(p["end_node"])
		ob, oe = int(p["own_begin"]), int(p["own_end"])

		# monotone local numbering, owned nodes contiguous and complete
		assert np.all(np.diff(l2g) > 0)
		assert np.array_equal(l2g[ob:oe], np.arange(lo, hi))
		assert np.all(l2g[:ob] < lo) and np.all(l2g[oe:] >= hi)

		# local elements = exactly the global elements touching an owned node, ascending, same connectivity
		touching = np.nonzero(((elems >= lo) & (elems < hi)).any(axis=1))[0]
		assert np.array_equal(p["elem_l2g"].astype(np.int64), touching)
		assert np.array_equal(l2g[p["elems"].astype(np.int64)], elems[touching])

		# ghosts = the other nodes of those elements
		ghosts = np.setdiff1d(np.unique(elems[touching]), np.arange(lo, hi))
		assert np.array_equal(np.concatenate([l2g[:ob], l2g[oe:]]), ghosts)

		# receive ranges: contiguous per neighbour, cover all ghosts, owner is right
		covered = np.zeros(len(l2g), bool)
		covered[ob:oe] = True

		for i, s in enumerate(p["nbr"]):
			beg, cnt = int(p["recv_begin"][i]), int(p["recv_count"][i])
			assert cnt > 0 and not covered[beg:beg + cnt].any()
			covered[beg:beg + cnt] = True
			assert np.all((l2g[beg:beg + cnt] >= parts[s]["first_node"]) & (l2g[beg:beg + cnt] < parts[s]["end_node"]))

		assert covered.all()
		assert np.all(np.diff(p["nbr"]) > 0) and r not in p["nbr"]

	# the two ends of every link agree on contents and order
	for r, p in enumerate(parts):
		for i, s in enumerate(p["nbr"]):
			q = parts[s]
			assert r in q["nbr"]
			j = list(q["nbr"]).index(r)
			sent = p["l2g"][p["send_idx"][p["send_ptr"][i]:p["send_ptr"][i + 1]]]
			expected = q["l2g"][q["recv_begin"][j]:q["recv_begin"][j] + q["recv_count"][j]]
			assert np.array_equal(sent, expected)

	return parts


@pytest.mark.parametrize("name", MESHES)
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_structure(name, world, lib):
	case = cases.build(name, lib)
	_check_partition(case.mesh, world)


def test_partition_rejects_more_ranks_than_nodes(lib):
	mesh = api.Mesh.from_arrays(np.array([[0.0, 0], [1, 0], [0, 1]]), np.array([[0, 1, 2]]), binding=lib)
	info = ext.PartitionInfo()

	os.environ["BFM_QUIET"] = "1"

	try:
		assert lib.lib.bfmx_partition_sizes(mesh.c_mesh, 0, 4, info) == -1
	finally:
		del os.environ["BFM_QUIET"]


@pytest.mark.parametrize("name,world", [("plate_40x10", 2), ("plate_40x10", 4), ("bridge", 3), ("plate_q4_24x6", 2)])
def test_emulated_partitioned_pcg_matches_oracle(name, world, lib, golden):
	"""all ranks emulated in one process: owned rows + halo plan are enough to solve to the reference's answer"""

	case = cases.build(name, lib)
	oracle = cases.oracle_problem(case).system()
	A, b = oracle.scipy(), oracle.b.copy()

	views = [emu.RankView(ext.partition(case.mesh, r, world), A, b) for r in range(world)]

	def exchange(vectors):
		packed = {(r, int(s)): views[r].pack(vectors[r], i) for r in range(world) for i, s in enumerate(views[r].part["nbr"])}

		for r in range(world):
			for i, s in enumerate(views[r].part["nbr"]):
				views[r].unpack(vectors[r], i, packed[(int(s), r)])

	xs, iters = emu.pcg(views, exchange, lambda parts: float(sum(parts)))
	x = np.concatenate(xs)

	want = golden[f"{name}/effects"].reshape(-1)
	assert np.linalg.norm(x - want) / np.linalg.norm(want) <= 1e-9, iters


def test_two_rank_gloo_pcg(tmp_path):
	"""one process per rank, torch.distributed gloo, world_size 2: halo via send/recv, dots via all_reduce"""

	out = tmp_path / "result.txt"
	cmd = [
		sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
		"--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
		os.path.join(ROOT, "tests", "dist_worker.py"), "plate_40x10", str(out),
	]

	proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
	assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]

	err, iters = out.read_text().split()
	assert float(err) <= 1e-9 and int(iters) > 10
