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


def test_emulated_partitioned_two_level_pcg(lib, golden):
	"""the distributed form of the coarse level (coarse.cuh): every rank restricts over its OWNED rows only, the
	partial W^T r are summed over the ranks, E^-1 is applied, and each rank prolongs to its own rows - with the
	aggregates the library computes on the global mesh.  Same answer as the reference, far fewer iterations."""

	import ctypes as C

	import scipy.sparse as sp

	name, world = "plate_80x20", 3
	case = cases.build(name, lib)
	oracle = cases.oracle_problem(case).system()
	A, b = oracle.scipy(), oracle.b.copy()
	nn = case.mesh.n_nodes
	coords = case.mesh.coords_array

	n_agg, n_colors = C.c_int32(), C.c_int32()
	agg = np.zeros(nn, np.int32)
	color = np.zeros(4 * 40 + 16, np.int32)
	assert not lib.lib.bfmx_coarse_plan(case.mesh.c_mesh, 40, C.byref(n_agg), C.byref(n_colors), agg.ctypes.data_as(ext.c_int32_p), color.ctypes.data_as(ext.c_int32_p))
	n_agg = n_agg.value
	assert n_agg >= 16

	# W in the Jacobi-scaled variables: rows S_a R_a (coarse.cuh)
	d = 1 / np.sqrt(np.abs(A.diagonal()))
	Ah = (sp.diags(d) @ A @ sp.diags(d)).tocsr()
	bh = d * b
	count = np.bincount(agg, minlength=n_agg)
	cx = np.bincount(agg, coords[:, 0], n_agg) / count
	cy = np.bincount(agg, coords[:, 1], n_agg) / count
	a = np.arange(nn)
	w = 1 / d
	W = sp.csr_matrix((
		np.concatenate([w[2 * a], w[2 * a + 1], -w[2 * a] * (coords[:, 1] - cy[agg]), w[2 * a + 1] * (coords[:, 0] - cx[agg])]),
		(np.concatenate([2 * a, 2 * a + 1, 2 * a, 2 * a + 1]), np.concatenate([3 * agg, 3 * agg + 1, 3 * agg + 2, 3 * agg + 2])),
	), shape=(2 * nn, 3 * n_agg))
	Einv = np.linalg.inv((W.T @ Ah @ W).toarray())

	parts = [ext.partition(case.mesh, r, world) for r in range(world)]
	rows = [np.arange(2 * int(p["first_node"]), 2 * int(p["end_node"])) for p in parts]

	def precondition(r_full):
		g = sum(W[rows[k]].T @ r_full[rows[k]] for k in range(world))  # per-rank partial restrictions, folded in rank order
		mu = Einv @ g
		z = r_full.copy()

		for k in range(world):
			z[rows[k]] += W[rows[k]] @ mu  # each rank prolongs to its own rows

		return z

	def pcg(M):
		x = np.zeros_like(bh)
		r = bh.copy()
		z = M(r)
		p = z.copy()
		rz = r @ z

		for it in range(1, 5000):
			q = np.concatenate([Ah[rows[k]] @ p for k in range(world)])  # owned rows of the SpMV
			alpha = rz / (p @ q)
			x += alpha * p
			r -= alpha * q

			if r @ r <= 1e-24 * (bh @ bh):
				return d * x, it

			z = M(r)
			new = r @ z
			p = z + (new / rz) * p
			rz = new

		raise AssertionError("no convergence")

	x_two, it_two = pcg(precondition)
	_, it_plain = pcg(lambda r: r)

	want = golden[f"{name}/effects"].reshape(-1)

	assert np.linalg.norm(x_two - want) / np.linalg.norm(want) <= 1e-9
	assert it_two * 3 < it_plain, (it_two, it_plain)
