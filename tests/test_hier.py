"""Host side of the multilevel preconditioner (bfm_b200/csrc/hier.c), on CPU: structural invariants of the
aggregation hierarchy - every connected node has an aggregate, aggregates are big enough for three independent
rigid-body modes, every level's pattern is the symbolic product P^T A P (what k_mg_rap fills in) - and, through tests/mg_emulation.py, the numerics of the cycle the device runs on it."""

import numpy as np
import pytest
import scipy.sparse as sp

import cases
import mg_emulation


def _pattern(level):
	n = level["n"]
	return sp.csr_matrix((np.ones(len(level["col"]), np.int8), level["col"].astype(np.int64), level["rowptr"].astype(np.int64)), shape=(n, n))


@pytest.mark.parametrize("smooth", ["0", "1"])
@pytest.mark.parametrize("name", ["gear60", "plate_160x40", "plate_300x75", "plate_jitter_40x10", "bridge_dam"])
def test_hierarchy_invariants(name, smooth, lib, monkeypatch):
	monkeypatch.setenv("BFM_MG_SMOOTH", smooth)

	if name == "bridge_dam":
		monkeypatch.setenv("BFM_MG_RATIO0", "6")      # a small truss mesh: small aggregates so that it still gets levels
		monkeypatch.setenv("BFM_MG_DENSE_NODES", "64")

	case = cases.build(name, lib)
	levels = mg_emulation.hierarchy(lib, case.mesh)

	assert len(levels) >= 2 and levels[0]["n"] == case.mesh.n_nodes

	for l, (fine, coarse) in enumerate(zip(levels[:-1], levels[1:])):
		agg = fine["agg"]
		sizes = np.bincount(agg[agg >= 0], minlength=coarse["n"])

		assert agg.max() == coarse["n"] - 1 and sizes.min() >= (3 if l == 0 else 1)
		assert coarse["n"] * 10 <= fine["n"] * 7                     # it coarsens
		assert np.all(np.diff(np.unique(agg[agg >= 0], return_index=True)[1]) > 0)  # ids follow the smallest member

		# nodes left out are isolated (coupled to nothing but themselves)
		degree = np.diff(fine["rowptr"])
		assert np.all(degree[agg < 0] <= 1)

		# pattern of the next level == pattern of P^T A P, P the aggregate indicator - or, with smoothed aggregation,
		# the pattern of (I + A) times it: every node reaches the aggregates of its neighbours
		keep = np.flatnonzero(agg >= 0)
		P = sp.csr_matrix((np.ones(len(keep), np.int32), (keep, agg[keep].astype(np.int64))), shape=(fine["n"], coarse["n"]))

		if smooth == "1":
			mask = sp.diags((agg >= 0).astype(np.float64))
			P = (mask @ (_pattern(fine).astype(np.int32) @ P + P)).tocsr()
			P.data[:] = 1

		want = (P.T @ _pattern(fine).astype(np.int32) @ P).tocsr()
		want.sort_indices()
		got = _pattern(coarse)

		assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
		assert np.all(np.diff(coarse["col"].astype(np.int64))[np.setdiff1d(np.arange(len(coarse["col"]) - 1), coarse["rowptr"][1:-1] - 1)] > 0)  # ascending columns

		# geometry: relative to the centroid of the aggregate
		for k in range(2):
			sums = np.bincount(agg[keep], fine["geom"][keep, k], coarse["n"])
			scale = np.abs(fine["geom"]).max() + 1e-300
			assert np.abs(sums / sizes).max() <= 1e-5 * scale

	# deterministic
	again = mg_emulation.hierarchy(lib, case.mesh)
	assert all(np.array_equal(a["agg"], b["agg"]) and np.array_equal(a["col"], b["col"]) for a, b in zip(levels, again))


def test_no_hierarchy_for_tiny_meshes(lib):
	from bfm_b200 import ext

	assert mg_emulation.hierarchy(lib, ext.plate(4, 2, kind=3, binding=lib)) == []   # 15 nodes: nothing to coarsen
	assert len(mg_emulation.hierarchy(lib, cases.build("lepl8", lib).mesh)) == 2     # 335 nodes: mesh level + dense level


@pytest.mark.parametrize("name,most", [("gear60", 200), ("plate_160x40", 60), ("plate_300x75", 70)])
def test_emulated_smoothed_aggregation_converges_to_the_reference(name, most, lib, golden, monkeypatch):
	"""the same with smoothed aggregation (BFM_MG_SMOOTH=1: k_mg_smooth, V-cycle), on the hierarchy hier.c lays out for it"""

	monkeypatch.setenv("BFM_MG_SMOOTH", "1")

	case = cases.build(name, lib)
	levels = mg_emulation.hierarchy(lib, case.mesh)
	system = cases.oracle_problem(case).system()

	emu = mg_emulation.Emulation(system.scipy().tocsr(), levels, smooth=True)
	x, iterations = emu.solve(system.b.copy())

	assert iterations <= most, iterations

	want = golden[f"{name}/effects"].reshape(-1)
	assert np.linalg.norm(x - want) / np.linalg.norm(want) <= 1e-9


@pytest.mark.parametrize("name,most", [("gear60", 320), ("plate_160x40", 80), ("plate_jitter_40x10", 60)])
def test_emulated_cycle_converges_to_the_reference(name, most, lib, golden):
	"""the W-cycle of mg.cuh as a PCG preconditioner, emulated in numpy on the library's own hierarchy: converges
	to 1e-12 in a mesh-independent handful of iterations and lands on the reference's displacements"""

	case = cases.build(name, lib)
	levels = mg_emulation.hierarchy(lib, case.mesh)
	system = cases.oracle_problem(case).system()

	emu = mg_emulation.Emulation(system.scipy().tocsr(), levels)
	x, iterations = emu.solve(system.b.copy())

	assert iterations <= most, iterations
	assert all(0 < w * 1.0 < 1.6 for w in emu.omega)

	want = golden[f"{name}/effects"].reshape(-1)
	assert np.linalg.norm(x - want) / np.linalg.norm(want) <= 1e-9


def test_hierarchy_info_counts_prolongator_entries(lib, monkeypatch):
	"""bfmx_hier_info: one entry per node with the tentative prolongator, one per (node, aggregate of a neighbour) with
	smoothed aggregation - the count bench.py's byte model uses"""

	import ctypes as C

	from bfm_b200 import ext

	case = cases.build("plate_160x40", lib)
	counts = {}

	for smooth in ("0", "1"):
		monkeypatch.setenv("BFM_MG_SMOOTH", smooth)
		info = ext.HierInfo()
		assert not lib.lib.bfmx_hier_info(case.mesh.c_mesh, C.byref(info))
		assert info.smoothed == int(smooth) and info.n_levels >= 2 and info.n_entries[info.n_levels - 1] == 0
		counts[smooth] = (info.n_nodes[0], info.n_entries[0])

	levels = mg_emulation.hierarchy(lib, case.mesh)  # smoothed layout
	agg, rowptr, col = levels[0]["agg"], levels[0]["rowptr"], levels[0]["col"]
	want = sum(len(np.unique(agg[col[rowptr[a]:rowptr[a + 1]]])) for a in range(levels[0]["n"]))

	assert counts["0"] == (case.mesh.n_nodes, case.mesh.n_nodes)
	assert counts["1"] == (case.mesh.n_nodes, want) and want > 2 * case.mesh.n_nodes


def test_emulated_refinement_window(lib, monkeypatch):
	"""solver.cu's two ways below FP64's residual floor, in the numpy twin on a 0.18 M-DOF plate (at 0.5 M DOF the same
	experiment gives 2.7e-9 / 5.9e-15 / 1.3e-14 in 59 / 67 / 59 iterations): plain PCG stops 1e-10 short of the exact
	solution however small its recursive residual; the restart at the floor (phi = 6) gets to 1e-15 and pays iterations for
	the lost Krylov space; replacing the residual while KEEPING the search direction is free at phi = 1e4 - and stalls CG
	when done at the floor, which is why the device does it early and only once"""

	import scipy.sparse.linalg as spla

	monkeypatch.setenv("BFM_MG_SMOOTH", "1")

	case = cases.build("plate_600x150", lib)
	levels = mg_emulation.hierarchy(lib, case.mesh)
	system = cases.oracle_problem(case).system()
	A, b = system.scipy().tocsr(), system.b.copy()

	lu = spla.splu(A.tocsc())
	exact = lu.solve(b)

	for _ in range(3):
		exact = exact + lu.solve(mg_emulation.residual_extended(A, b, exact))

	emu = mg_emulation.Emulation(A, levels, smooth=True)
	err = lambda x: float(np.linalg.norm(x - exact) / np.linalg.norm(exact))

	x, plain_its, _ = mg_emulation.solve_refined(emu, A, b, "none", 0)
	assert 1e-11 < err(x) < 1e-8          # the floor

	x, restart_its, events = mg_emulation.solve_refined(emu, A, b, "restart", 6)
	assert err(x) < 1e-13 and events == 1 and restart_its > plain_its

	x, replace_its, events = mg_emulation.solve_refined(emu, A, b, "replace", 1e4, max_events=1)
	assert err(x) < 1e-13 and events == 1 and replace_its <= plain_its + 1

	x, stalled_its, _ = mg_emulation.solve_refined(emu, A, b, "replace", 6, max_iter=150, max_events=1)
	assert stalled_its >= 2 * plain_its
