"""The symbolic phase on the device (bfm_b200/csrc/symbolic.cu) against its host twins, array for array.

- sparsity plan: slice offsets, row lengths, columns, diagonal slots, contributor bounds and the packed
  (element, j, k) contributor lists must equal what plan.c's OpenMP builder produces (BFM_PLAN=host);
- edges: the list must equal the host merge sort's (BFM_EDGES=host), which tests/test_host.py pins to the
  reference's own reader (reference mesh.c:32-102) - order, node pairs and element pairs.

Integer work: the bar is equality.
"""

import ctypes as C
import os

import numpy as np
import pytest

import cases
from bfm_b200 import api, ext

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_device(lib):
	assert ext.device_available(lib), lib.lib.bfmx_device_error()


class _env:
	def __init__(self, **kv):
		self.kv = kv

	def __enter__(self):
		self.old = {k: os.environ.get(k) for k in self.kv}
		os.environ.update(self.kv)

	def __exit__(self, *exc):
		for k, v in self.old.items():
			if v is None:
				os.environ.pop(k, None)
			else:
				os.environ[k] = v


def fan_arrays(n: int):
	"""n triangles around node 0 (valence n: 3 n keys in one row - the heap-sort path of the per-node sort), with
	two isolated nodes (lone diagonal blocks) in the middle of the numbering and at its end"""

	angle = 2 * np.pi * np.arange(n) / n
	ring = np.c_[np.cos(angle), np.sin(angle)]
	coords = np.vstack([[0.0, 0.0], ring[: n // 2], [5.0, 5.0], ring[n // 2 :], [6.0, 6.0]])
	number = np.r_[np.arange(1, n // 2 + 1), np.arange(n // 2 + 2, n + 2)]  # ring node -> mesh node
	elems = np.stack([np.zeros(n, np.int64), number, np.roll(number, -1)], axis=1)

	return coords, elems.astype(np.uint64)


def shuffled(coords, elems, seed=1):
	"""the same mesh with its nodes renumbered at random and its elements in random order, rotated at random"""

	rng = np.random.RandomState(seed)
	perm = rng.permutation(len(coords))
	new_coords = np.empty_like(coords)
	new_coords[perm] = coords
	new_elems = perm[elems.astype(np.int64)][rng.permutation(len(elems))]
	new_elems = np.stack([np.roll(row, rng.randint(elems.shape[1])) for row in new_elems])

	return new_coords, new_elems.astype(np.uint64)


def mesh_arrays(name, lib):
	if name.startswith("plate"):
		_, kind, nx, ny = name.split("_")
		return cases.plate_arrays(int(nx), int(ny), kind=int(kind))

	if name.startswith("shuffled"):
		_, kind, nx, ny = name.split("_")
		return shuffled(*cases.plate_arrays(int(nx), int(ny), kind=int(kind)))

	if name.startswith("fan"):
		return fan_arrays(int(name.split("_")[1]))

	mesh = cases.build(name, lib).mesh
	return mesh.coords_array.copy(), mesh.elems_array.copy()


MESHES = ["lepl8", "bridge", "gear60", "plate_3_33_7", "plate_4_24_6", "plate_3_1_1", "plate_3_700_300", "plate_4_500_200", "shuffled_3_300_100", "shuffled_4_90_70", "fan_7", "fan_40", "fan_500"]


@pytest.mark.parametrize("name", MESHES)
def test_device_plan_equals_host_plan(name, lib):
	coords, elems = mesh_arrays(name, lib)

	with _env(BFM_PLAN="host"):
		want = ext.pattern(api.Mesh.from_arrays(coords, elems, binding=lib))

	got = ext.pattern(api.Mesh.from_arrays(coords, elems, binding=lib))

	for key in ("n_slices", "n_slots", "n_blocks", "n_ctr", "nb"):
		assert got[key] == want[key], key

	for key in ("slice_off", "row_len", "scol", "diag_pos", "ctr_ptr", "ctr"):
		assert np.array_equal(got[key], want[key]), key


@pytest.mark.parametrize("name", MESHES)
def test_device_edges_equal_host_edges(name, lib):
	coords, elems = mesh_arrays(name, lib)

	def edges(how):
		mesh = api.Mesh.from_arrays(coords, elems, binding=lib)

		with _env(BFM_EDGES=how):
			rv = lib.lib.bfmx_mesh_compute_edges(C.byref(mesh.c_mesh))

		return rv, mesh.edges_array

	rv_host, want = edges("host")
	rv_dev, got = edges("device")

	assert rv_dev == rv_host == 0
	assert got.shape == want.shape and np.array_equal(got, want)


def test_device_edges_of_a_non_manifold_mesh(lib):
	"""three triangles on one edge and a duplicated element: the greedy fusion of the reference pairs neighbours of
	the sorted list only; the per-node walk must replay exactly that"""

	coords = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [0.5, -1], [2, 2]], np.float64)
	elems = np.array([[0, 1, 2], [1, 0, 4], [0, 1, 3], [1, 0, 4], [2, 1, 3], [0, 1, 2]], np.uint64)

	results = []

	for how in ("host", "device"):
		mesh = api.Mesh.from_arrays(coords, elems, binding=lib)

		with _env(BFM_EDGES=how):
			assert lib.lib.bfmx_mesh_compute_edges(C.byref(mesh.c_mesh)) == 0

		results.append(mesh.edges_array)

	assert np.array_equal(results[0], results[1])


def test_large_plate_runs_on_the_device_built_plan(lib):
	"""the plan the kernels built is the one bfm_sim_run uses: 0.5 M DOF, bit-identical displacements with either builder"""

	from bfm_b200 import workloads

	out = []

	for how in ("host", "device"):
		with _env(BFM_PLAN=how):
			case = workloads.plate_case(1000, 250, binding=lib)
			case.sim.run()

		stats = ext.last_stats(lib)
		assert stats["cg_converged"] == 1
		out.append(np.array(case.instance.effects))

	assert np.array_equal(out[0], out[1])


# ---- internal numbering (bfm_b200/csrc/renumber.c) ---------------------------------------------------------------


@pytest.mark.parametrize("name", ["lepl8_all_kinds", "plate_neumann_16x4", "plate_funky_20x5", "plate_q4_24x6", "lepl8_axisym", "bridge", "gear60", "plate_160x40"])
@pytest.mark.parametrize("one_cta", ["1", "0"])
def test_renumbered_run_matches_reference(name, one_cta, lib, golden):
	"""bfm_sim_run on the internally renumbered copy of the mesh (forced: these meshes are small and well numbered):
	all condition kinds, non-zero Dirichlet values, Neumann loads in edge order, FUNKY forces, quads, axisymmetry -
	displacements come back in the caller's numbering, within 1e-9 of the reference's direct solve"""

	with _env(BFM_RENUMBER="1", BFM_ONE_CTA=one_cta):
		case = cases.build(name, lib)
		assert ext.internal_numbering(case.mesh) is not None
		case.sim.run()

	stats = ext.last_stats(lib)
	want = golden[f"{name}/effects"]

	assert stats["cg_converged"] == 1
	assert float(np.linalg.norm(np.asarray(case.instance.effects) - want) / np.linalg.norm(want)) <= 1e-9


def test_randomly_numbered_plate_on_its_internal_numbering(lib):
	"""the plate with nodes numbered at random and elements in random order: solved on the Morton numbering it gives
	the displacements of the naturally numbered plate (compared node by node through the permutation)"""

	from bfm_b200 import workloads

	nx, ny = 600, 150
	coords, elems = cases.plate_arrays(nx, ny)
	new_coords, new_elems = shuffled(coords, elems, seed=5)

	# recover the permutation the helper drew: new_coords[perm] = coords
	perm = np.random.RandomState(5).permutation(len(coords))

	def solve(c, e, renumber):
		with _env(BFM_RENUMBER=renumber):
			mesh = api.Mesh.from_arrays(c, e, binding=lib)
			left = c[:, 0] == 0.0
			E, nu, rho = workloads.STEEL
			material = api.Material("steel", rho, E, nu, binding=lib)
			rule = api.Rule_gauss_legendre(2, mesh.kind, binding=lib)
			obj = api.Obj(mesh, material, rule)
			instance = api.Instance(obj)
			keep = [mesh, material, rule, obj]

			for kind in (api.Condition.DIRICHLET_X, api.Condition.DIRICHLET_Y):
				cond = api.Condition(mesh, kind, 0.0)
				cond.set_nodes(left)
				instance.add_condition(cond)
				keep.append(cond)

			sim = api.Sim(api.CSim.PLANAR_STRESS, binding=lib)
			sim.add_instance(instance)
			force = api.Force_linear(workloads.GRAVITY, binding=lib)
			sim.add_force(force)
			sim.run()

			assert ext.last_stats(lib)["cg_converged"] == 1
			return workloads.effects_view(instance).reshape(-1, 2).copy()

	natural = solve(coords, elems, "0")
	kept = solve(new_coords, new_elems, "0")
	morton = solve(new_coords, new_elems, "1")

	norm = np.linalg.norm(natural)

	assert np.linalg.norm(kept[perm] - natural) / norm <= 1e-9
	assert np.linalg.norm(morton[perm] - natural) / norm <= 1e-9
