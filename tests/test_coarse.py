"""Host side of the solver's coarse level (bfm_b200/csrc/coarse.c), on CPU: every node gets an aggregate,
aggregates are big enough to carry three independent rigid-body modes, and the colouring is a distance-2
colouring of the aggregate graph - the property the probing of E = W^T A W on the device relies on."""

import ctypes as C

import numpy as np
import pytest

import cases
from bfm_b200 import ext


def _plan(lib, mesh, target):
	n_agg, n_colors = C.c_int32(), C.c_int32()
	agg = np.full(mesh.n_nodes, -1, np.int32)
	color = np.full(4 * target + 16, -1, np.int32)

	assert not lib.lib.bfmx_coarse_plan(mesh.c_mesh, target, C.byref(n_agg), C.byref(n_colors), agg.ctypes.data_as(ext.c_int32_p), color.ctypes.data_as(ext.c_int32_p))

	return n_agg.value, n_colors.value, agg, color[:n_agg.value]


@pytest.mark.parametrize("name,target", [("bridge", 16), ("bridge_dam", 24), ("gear60", 64), ("plate_80x20", 16), ("plate_q4_24x6", 8), ("lepl8", 8)])
def test_aggregates_and_colouring(name, target, lib):
	case = cases.build(name, lib)
	n_agg, n_colors, agg, color = _plan(lib, case.mesh, target)

	assert 4 <= n_agg <= 4 * target
	assert agg.min() == 0 and agg.max() == n_agg - 1
	assert np.bincount(agg, minlength=n_agg).min() >= 3           # three independent modes per aggregate
	assert color.min() == 0 and color.max() == n_colors - 1

	# aggregate graph from the elements: g ~ h when an element has nodes in both
	adjacent = [set() for _ in range(n_agg)]

	for row in agg[case.mesh.elems_array.astype(np.int64)]:
		for g in row:
			adjacent[g].update(int(h) for h in row if h != g)

	# distance-2: an aggregate and all its neighbours carry pairwise different colours
	for h in range(n_agg):
		near = [h] + sorted(adjacent[h])
		assert len({int(color[g]) for g in near}) == len(near), (h, near)

	# deterministic
	again = _plan(lib, case.mesh, target)
	assert np.array_equal(again[2], agg) and np.array_equal(again[3], color)


def test_no_coarse_level_for_tiny_or_degenerate_meshes(lib):
	case = cases.build("plate_q4_24x6", lib)

	assert _plan(lib, case.mesh, 2)[0] == 0            # fewer than 4 aggregates asked for
	assert _plan(lib, case.mesh, 10_000)[0] == 0       # more aggregates than nodes / 3
