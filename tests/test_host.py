"""CPU-side tests (no GPU): the C ABI surface, the host stages either side of the hot path (mesh readers,
problem parser, symbolic plan, RCM, dense/band matrices) and the loud failure when no device exists."""

import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
from bfm_b200 import _abi as abi
from bfm_b200 import api, ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- ABI -------------------------------------------------------------------------------------------


def _declared(header):
	text = open(os.path.join(ROOT, "include", header)).read()
	text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
	text = re.sub(r"//[^\n]*", "", text)
	return set(re.findall(r"\b(bfmx?_[a-z0-9_]+)\s*\(", text)) - {"bfm_alloc_t", "bfm_realloc_t", "bfm_free_t"}


def test_library_exports_every_declared_symbol(lib):
	raw = C.CDLL(api.LIB_PATH)
	core = {n for n in _declared("bfm/libbfm.h") if not n.endswith("_t")}
	extra = {n for n in _declared("bfm_b200.h") if not n.endswith("_t")}

	assert core == set(abi.PROTOTYPES), core ^ set(abi.PROTOTYPES)  # the reference's 67 functions
	assert len(core) == 67

	for name in sorted(core | extra):
		assert hasattr(raw, name), f"{name} declared in include/ but not exported"


def test_forwarding_headers_compile(tmp_path):
	"""sources written against the reference include <bfm/sim.h> etc.: every forwarder must work alone"""

	import subprocess

	names = "bfm math matrix mesh condition force material shape rule obj instance sim perm system ez".split()

	for name in names:
		src = tmp_path / f"use_{name}.c"
		src.write_text(f"#include <bfm/{name}.h>\nint main(void) {{ return (int) sizeof(bfm_state_t) - 64; }}\n")
		subprocess.run(["gcc", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


def test_struct_sizes_match_the_c_compiler(tmp_path):
	import subprocess

	types = {cls.__name__: size for cls, size in abi.EXPECTED_SIZES.items()}
	c_names = {
		"State": "bfm_state_t", "Vec": "bfm_vec_t", "Matrix": "bfm_matrix_t", "Perm": "bfm_perm_t", "System": "bfm_system_t",
		"Mesh": "bfm_mesh_t", "Edge": "bfm_edge_t", "Domain": "bfm_domain_t", "Condition": "bfm_condition_t", "Force": "bfm_force_t",
		"Material": "bfm_material_t", "Shape": "bfm_shape_t", "Rule": "bfm_rule_t", "Obj": "bfm_obj_t", "Instance": "bfm_instance_t",
		"Sim": "bfm_sim_t", "Ez": "bfm_ez_lepl1110_t",
	}

	checks = "\n".join(f'_Static_assert(sizeof({c_names[n]}) == {s}, "{n}");' for n, s in types.items())
	src = tmp_path / "sizes.c"
	src.write_text(f"#include <bfm/libbfm.h>\n{checks}\nint main(void) {{ return 0; }}\n")
	subprocess.run(["gcc", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


# ---- readers and parser ----------------------------------------------------------------------------


@pytest.mark.parametrize("name", ["8.lepl1110", "gear60_full.lepl1110"])
def test_lepl_reader(name, lib):
	path = os.path.join(cases.GOLDEN, "meshes", name)
	mesh = api.Mesh_lepl1110(path, binding=lib)
	coords, elems, edges, domains = cases._read_lepl(path)

	assert np.array_equal(mesh.coords_array, coords)
	assert np.array_equal(mesh.elems_array, elems)
	assert np.array_equal(mesh.edges_array, edges)
	assert list(mesh.domains()) == list(domains)

	for got, want in zip(mesh.domains().values(), domains.values()):
		assert np.array_equal(got, np.array(want, dtype=np.uint64))


def test_lepl_reader_rejects_files_without_edge_and_domain_sections(lib, tmp_path):
	"""meshes/gear60.lepl1110 as shipped by the reference: its own reader returns -1"""

	text = open(os.path.join(cases.GOLDEN, "meshes", "8.lepl1110")).read()
	start = text.index("Number of edges")
	stop = text.index("Number of quads")
	broken = tmp_path / "no_edges.lepl1110"
	broken.write_text(text[:start] + text[stop:text.index("Number of domains")])

	mesh = abi.Mesh()

	assert lib.lib.bfm_mesh_read_lepl1110(C.byref(mesh), C.byref(lib.state), str(broken).encode()) == -1
	assert lib.lib.bfm_mesh_read_lepl1110(C.byref(mesh), C.byref(lib.state), b"/nonexistent") == -1


@pytest.mark.parametrize("name", ["8.lepl1110"])
def test_readers_match_live_reference(name, lib, ref):
	path = os.path.join(cases.GOLDEN, "meshes", name)
	ours, theirs = api.Mesh_lepl1110(path, binding=lib), api.Mesh_lepl1110(path, binding=ref)

	assert np.array_equal(ours.coords_array, theirs.coords_array)
	assert np.array_equal(ours.elems_array, theirs.elems_array)
	assert np.array_equal(ours.edges_array, theirs.edges_array)


@pytest.mark.parametrize("name", ["bridge.obj", "bridge-dam.obj"])
def test_wavefront_reader_matches_live_reference(name, lib, ref):
	path = os.path.join(cases.GOLDEN, "meshes", name)

	for full in (False, True):
		ours, theirs = api.Mesh_wavefront(path, full, binding=lib), api.Mesh_wavefront(path, full, binding=ref)

		assert ours.c_mesh.dim == theirs.c_mesh.dim == (3 if full else 2)
		assert np.array_equal(ours.coords_array, theirs.coords_array)
		assert np.array_equal(ours.elems_array, theirs.elems_array)
		assert np.array_equal(ours.edges_array, theirs.edges_array)  # order matters: Neumann loads follow it


def test_wavefront_edges_invariants(lib):
	mesh = api.Mesh_wavefront(os.path.join(cases.GOLDEN, "meshes", "bridge.obj"), binding=lib)
	edges = mesh.edges_array

	# Euler: V - E + F = 1 - holes for a planar triangulation; every edge has 1 or 2 elements
	assert (edges[:, 2] >= 0).all() and ((edges[:, 3] >= 0) | (edges[:, 3] == -1)).all()

	lo = np.minimum(edges[:, 0], edges[:, 1])

	assert (np.diff(lo) <= 0).all()  # sorted by smaller node, descending (reference mesh.c:32-50)
	assert len({(min(a, b), max(a, b)) for a, b in edges[:, :2]}) == len(edges)


def test_plate_generator_matches_numpy_twin(lib):
	for kind in (3, 4):
		mesh = ext.plate(13, 7, kind=kind, binding=lib)
		coords, elems = cases.plate_arrays(13, 7, kind=kind)

		assert np.array_equal(mesh.coords_array, coords)
		assert np.array_equal(mesh.elems_array, elems)


@pytest.mark.parametrize("problem", ["problem.txt", "lepl8_all_kinds.txt", "lepl8_axisym.txt"])
def test_problem_parser(problem, lib):
	mesh_path = os.path.join(cases.GOLDEN, "meshes", "8.lepl1110")
	mesh = api.Mesh_lepl1110(mesh_path, binding=lib)
	ez = api.Ez_lepl1110(mesh, os.path.join(cases.GOLDEN, "problems", problem))

	coords, elems, edges, domains = cases._read_lepl(mesh_path)
	sim_kind, E, nu, rho, g, conds = cases._read_problem(os.path.join(cases.GOLDEN, "problems", problem), len(coords), edges, domains)

	assert ez.sim.kind == sim_kind
	assert (ez.c_ez.material.E, ez.c_ez.material.nu, ez.c_ez.material.rho) == (E, nu, rho)
	assert ez.c_ez.gravity.linear.force.data[0] == 0 and ez.c_ez.gravity.linear.force.data[1] == g
	assert ez.c_ez.sim.n_forces == 1 and ez.c_ez.sim.n_instances == 1
	assert len(ez.conditions()) == len(conds) == ez.c_ez.instance.n_conditions

	for (k1, v1, m1), (k2, v2, m2) in zip(ez.conditions(), conds):
		assert (k1, v1) == (k2, v2) and np.array_equal(m1, m2) and m1.sum() > 0


def test_problem_parser_matches_live_reference(lib, ref):
	for binding in (lib, ref):
		mesh = api.Mesh_lepl1110(os.path.join(cases.GOLDEN, "meshes", "8.lepl1110"), binding=binding)
		ez = api.Ez_lepl1110(mesh, os.path.join(cases.GOLDEN, "problems", "lepl8_all_kinds.txt"))
		got = [(k, v, m.tobytes()) for k, v, m in ez.conditions()] + [ez.sim.kind, ez.c_ez.material.E]

		if binding is lib:
			ours = got

	assert ours == got


def test_uv_writer_format(lib, tmp_path):
	mesh = api.Mesh_lepl1110(os.path.join(cases.GOLDEN, "meshes", "8.lepl1110"), binding=lib)
	ez = api.Ez_lepl1110(mesh, os.path.join(cases.GOLDEN, "problems", "problem.txt"))

	golden = cases.golden()["lepl8/effects"]
	C.memmove(ez.c_ez.instance.effects, golden.ctypes.data, golden.nbytes)  # the reference's displacements

	for shift, name in ((0, "lepl8_U.txt"), (1, "lepl8_V.txt")):
		out = tmp_path / name
		ez.write(str(out), shift)

		assert out.read_text() == open(os.path.join(cases.GOLDEN, name)).read()  # byte for byte


# ---- symbolic plan ---------------------------------------------------------------------------------


def _element_terms(coords, elem, kind, weights, points, a, b, c):
	"""per integration point: (det*w, dphi_dx[], dphi_dy[]) in the reference's arithmetic (system.c:136-186)"""

	x, y = coords[elem, 0], coords[elem, 1]
	out = []

	for g in range(len(weights)):
		xsi, eta = points[g]

		if kind == 3:
			dxsi, deta = [-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]
		else:
			dxsi = [(1 + eta) / 4, (-1 - eta) / 4, (-1 + eta) / 4, (1 - eta) / 4]
			deta = [(1 + xsi) / 4, (1 - xsi) / 4, (-1 + xsi) / 4, (-1 - xsi) / 4]

		dx_dxsi = dx_deta = dy_dxsi = dy_deta = 0.0

		for j in range(kind):
			dx_dxsi += x[j] * dxsi[j]
			dx_deta += x[j] * deta[j]
			dy_dxsi += y[j] * dxsi[j]
			dy_deta += y[j] * deta[j]

		det = abs(dx_dxsi * dy_deta - dx_deta * dy_dxsi)
		dpx = [(dxsi[j] * dy_deta - deta[j] * dy_dxsi) / det for j in range(kind)]
		dpy = [(deta[j] * dx_dxsi - dxsi[j] * dx_deta) / det for j in range(kind)]
		out.append((det * weights[g], dpx, dpy))

	return out


@pytest.mark.parametrize("name", ["plate_q4_24x6", "plate_funky_20x5", "lepl8"])
def test_plan_pattern_and_contributor_map(name, lib):
	"""walking the element-to-nonzero map in list order reproduces the oracle's matrix bit for bit -
	the exact job the assembly kernel does on the GPU (a pure-Python emulation of assembly.cu)"""

	from oracle import orc

	problem = cases.build_oracle_only(name)
	oracle = problem.system(with_bcs=False)
	mesh = api.Mesh.from_arrays(problem.coords, problem.elems, binding=lib)
	P = ext.pattern(mesh)
	kind = problem.kind
	weights, points = orc.gauss_legendre(kind)

	E, nu = problem.c.E, problem.c.nu
	stress = problem.c.sim_kind == 2
	a = E / (1 - nu * nu) if stress else E * (1 - nu) / (1 + nu) / (1 - 2 * nu)
	b = E * nu / (1 - nu * nu) if stress else E * nu / (1 + nu) / (1 - 2 * nu)
	c = E / (2 * (1 + nu))

	assert P["n_ctr"] == len(problem.elems) * kind * kind
	assert P["n_slots"] % 32 == 0 and P["n_blocks"] == int(P["row_len"].sum())

	elems = problem.elems.astype(np.int64)
	cache = {}

	for node in range(P["nb"]):
		cols = []

		for t in range(P["row_len"][node]):
			slot = P["slice_off"][node // 32] + 32 * t + node % 32
			col = int(P["scol"][slot])
			cols.append(col)

			if col == node:
				assert P["diag_pos"][node] == slot

			acc = [0.0, 0.0, 0.0, 0.0]
			last = (-1, -1, -1)

			for packed in P["ctr"][P["ctr_ptr"][slot]:P["ctr_ptr"][slot + 1]]:
				e, j, k = int(packed) >> 4, (int(packed) >> 2) & 3, int(packed) & 3

				assert (e, j, k) > last and elems[e, j] == node and elems[e, k] == col
				last = (e, j, k)

				if e not in cache:
					cache[e] = _element_terms(problem.coords, elems[e], kind, weights, points, a, b, c)

				for dw, dpx, dpy in cache[e]:
					acc[0] += dw * (a * dpx[j] * dpx[k] + c * dpy[j] * dpy[k])
					acc[1] += dw * (b * dpx[j] * dpy[k] + c * dpy[j] * dpx[k])
					acc[2] += dw * (b * dpy[j] * dpx[k] + c * dpx[j] * dpy[k])
					acc[3] += dw * (a * dpy[j] * dpy[k] + c * dpx[j] * dpx[k])

			for r in range(2):
				row = 2 * node + r
				at = int(oracle.rowptr[row]) + 2 * t

				assert oracle.col[at] == 2 * col and oracle.col[at + 1] == 2 * col + 1
				assert oracle.val[at] == acc[2 * r] and oracle.val[at + 1] == acc[2 * r + 1]

		assert cols == sorted(set(cols))
		assert int(oracle.rowptr[2 * node + 1] - oracle.rowptr[2 * node]) == 2 * len(cols)

	# padding slots point at their own row and have no contributors
	for s in range(P["n_slices"]):
		for slot in range(P["slice_off"][s], P["slice_off"][s + 1]):
			row = 32 * s + (slot - P["slice_off"][s]) % 32
			t = (slot - P["slice_off"][s]) // 32

			if row >= P["nb"] or t >= P["row_len"][row]:
				assert P["ctr_ptr"][slot] == P["ctr_ptr"][slot + 1]
				assert P["scol"][slot] == min(row, P["nb"] - 1)


def test_plan_handles_isolated_nodes_and_is_cached(lib):
	coords = np.array([[0, 0], [1, 0], [0, 1], [5, 5]], dtype=np.float64)  # node 3 belongs to no element
	mesh = api.Mesh.from_arrays(coords, np.array([[0, 1, 2]], dtype=np.uint64), binding=lib)
	P = ext.pattern(mesh)

	assert list(P["row_len"]) == [3, 3, 3, 1]
	assert P["scol"][P["diag_pos"][3]] == 3 and P["n_ctr"] == 9

	Q = ext.pattern(mesh)

	assert all(np.array_equal(P[k], Q[k]) for k in ("scol", "ctr", "ctr_ptr"))

	bad = api.Mesh.from_arrays(coords, np.array([[0, 1, 9]], dtype=np.uint64), binding=lib)
	sizes = [C.c_size_t() for _ in range(4)]

	os.environ["BFM_QUIET"] = "1"
	assert lib.lib.bfmx_mesh_pattern_sizes(C.byref(bad.c_mesh), *[C.byref(s) for s in sizes]) == -1
	del os.environ["BFM_QUIET"]


# ---- RCM, dense and band matrices --------------------------------------------------------------------


def _full_system(lib, A, b):
	system = abi.System()
	n = len(b)

	assert not lib.lib.bfm_system_create(C.byref(system), C.byref(lib.state), n)
	assert system.A.kind == abi.MATRIX_KIND_FULL

	A = np.ascontiguousarray(A)
	C.memmove(system.A.full.data, A.ctypes.data, A.nbytes)
	C.memmove(system.b.data, b.ctypes.data, b.nbytes)

	return system


@pytest.mark.parametrize("name", ["lepl8", "plate_neumann_16x4", "plate_40x10"])
def test_dense_path_reproduces_the_reference(name, lib, golden):
	"""the public FULL/BAND API: bfm_system_create + renumber + bfm_matrix_solve, bit for bit"""

	oracle = cases.build_oracle_only(name).system()
	system = _full_system(lib, oracle.dense(), oracle.b.copy())
	n = oracle.n

	assert lib.lib.bfm_matrix_bandwidth(C.byref(system.A)) == int(golden[f"{name}/bandwidth_natural"])
	assert not lib.lib.bfm_system_renumber(C.byref(system))
	assert system.A.kind == abi.MATRIX_KIND_BAND and system.A.band.k == int(golden[f"{name}/bandwidth_rcm"])

	perm = np.ctypeslib.as_array(system.perm.perm, shape=(n,)).astype(np.int64)

	assert np.array_equal(perm, golden[f"{name}/perm"])

	assert not lib.lib.bfm_matrix_solve(C.byref(system.A), C.byref(system.b))
	assert not lib.lib.bfm_perm_perm_vec(C.byref(system.perm), C.byref(system.b), True)

	x = np.ctypeslib.as_array(system.b.data, shape=(n,)).reshape(-1, 2)

	assert np.array_equal(x, golden[f"{name}/effects"])

	lib.lib.bfm_system_destroy(C.byref(system))


def test_full_lu_against_numpy(lib):
	rng = np.random.default_rng(1)
	n = 40
	A = rng.standard_normal((n, n)) + n * np.eye(n)
	b = rng.standard_normal(n)

	for major in (abi.MATRIX_MAJOR_ROW, abi.MATRIX_MAJOR_COLUMN):
		M = abi.Matrix()
		assert not lib.lib.bfm_matrix_full_create(C.byref(M), C.byref(lib.state), major, n)

		for i in range(n):
			for j in range(n):
				assert not lib.lib.bfm_matrix_set(C.byref(M), i, j, A[i, j])

		assert lib.lib.bfm_matrix_get(C.byref(M), 3, 7) == A[3, 7]
		assert np.isnan(lib.lib.bfm_matrix_get(C.byref(M), n, 0))
		assert lib.lib.bfm_matrix_set(C.byref(M), 0, n, 1.0) == -1

		v = api.Vec(b.tolist(), lib)

		assert not lib.lib.bfm_matrix_solve(C.byref(M), C.byref(v.c_vec))
		assert np.allclose(np.ctypeslib.as_array(v.c_vec.data, shape=(n,)), np.linalg.solve(A, b), rtol=1e-10)

		lib.lib.bfm_matrix_destroy(C.byref(M))

	# a zero pivot is reported, not divided by
	Z = abi.Matrix()
	lib.lib.bfm_matrix_full_create(C.byref(Z), C.byref(lib.state), abi.MATRIX_MAJOR_ROW, 3)

	assert lib.lib.bfm_matrix_lu(C.byref(Z)) == -1

	lib.lib.bfm_matrix_destroy(C.byref(Z))


def test_band_storage_rules(lib):
	M = abi.Matrix()
	assert not lib.lib.bfm_matrix_band_create(C.byref(M), C.byref(lib.state), abi.MATRIX_MAJOR_ROW, 10, 2)

	assert lib.lib.bfm_matrix_bandwidth(C.byref(M)) == 2
	assert not lib.lib.bfm_matrix_set(C.byref(M), 4, 6, 3.5)
	assert lib.lib.bfm_matrix_get(C.byref(M), 4, 6) == 3.5
	assert lib.lib.bfm_matrix_get(C.byref(M), 4, 7) == 0.0          # outside the band reads as zero
	assert lib.lib.bfm_matrix_set(C.byref(M), 4, 7, 1e-30) == 0     # near-zero writes are swallowed
	assert lib.lib.bfm_matrix_set(C.byref(M), 4, 7, 1.0) == -1      # anything else is refused
	assert M.band.data[6 + 4 * 2 * 2] == 3.5                        # skewed addressing j + i * 2k

	lib.lib.bfm_matrix_destroy(C.byref(M))


@pytest.mark.parametrize("name", list(cases.CASES))
def test_sparse_rcm_is_bit_identical(name, lib, golden):
	"""bfm_perm_rcm on a CSR-kind matrix (host mirror only - no device needed)"""

	oracle = cases.build_oracle_only(name).system()
	matrix = ext.csr_matrix(lib, oracle.rowptr, oracle.col, oracle.val)
	perm = abi.Perm()

	assert not lib.lib.bfm_perm_create(C.byref(perm), C.byref(lib.state), oracle.n)
	assert lib.lib.bfm_matrix_bandwidth(C.byref(matrix)) == int(golden[f"{name}/bandwidth_natural"])
	assert not lib.lib.bfm_perm_rcm(C.byref(perm), C.byref(matrix))
	assert np.array_equal(np.ctypeslib.as_array(perm.perm, shape=(oracle.n,)).astype(np.int64), golden[f"{name}/perm"])

	assert not lib.lib.bfm_perm_perm_matrix(C.byref(perm), C.byref(matrix), False)
	assert lib.lib.bfm_matrix_bandwidth(C.byref(matrix)) == int(golden[f"{name}/bandwidth_rcm"])

	p = np.ctypeslib.as_array(perm.perm, shape=(oracle.n,))
	i = oracle.n // 3
	t = int(oracle.rowptr[i])

	assert lib.lib.bfm_matrix_get(C.byref(matrix), int(p[i]), int(p[int(oracle.col[t])])) == oracle.val[t]

	rowptr, col, val = ext.csr_export(lib, matrix)

	assert np.array_equal(rowptr, oracle.rowptr) and np.array_equal(col, oracle.col) and np.array_equal(val, oracle.val)

	lib.lib.bfm_perm_destroy(C.byref(perm))
	lib.lib.bfm_matrix_destroy(C.byref(matrix))


# ---- plumbing ------------------------------------------------------------------------------------------


def test_allocator_hooks_are_honoured(lib):
	state = abi.State()
	lib.lib.bfm_state_create(C.byref(state))

	libc = C.CDLL(None)
	libc.malloc.restype = C.c_void_p
	libc.malloc.argtypes = [C.c_size_t]
	libc.free.argtypes = [C.c_void_p]
	live = set()

	@abi.ALLOC_FN
	def counting_alloc(size):
		ptr = libc.malloc(size)
		live.add(ptr)
		return ptr

	@abi.FREE_FN
	def counting_free(ptr):
		live.discard(ptr)
		libc.free(ptr)

	assert not lib.lib.bfm_set_alloc(C.byref(state), counting_alloc)
	assert not lib.lib.bfm_set_free(C.byref(state), counting_free)

	vec = abi.Vec()

	assert not lib.lib.bfm_vec_create(C.byref(vec), C.byref(state), 16)
	assert len(live) == 1 and all(vec.data[i] == 0 for i in range(16))

	lib.lib.bfm_vec_destroy(C.byref(vec))

	assert not live


def test_force_eval_semantics(lib):
	force = api.Force_linear((1.5, -2.5), lib)
	pos, out = api.Vec((0, 0), lib), api.Vec((9, 9), lib)

	# the linear evaluator copies the vector and - like the reference (force.c:65-73) - still returns -1
	assert lib.lib.bfm_force_eval(C.byref(force.c_force), C.byref(pos.c_vec), C.byref(out.c_vec)) == -1
	assert (out.c_vec.data[0], out.c_vec.data[1]) == (1.5, -2.5)

	none = api.Force_none(2, lib)

	assert lib.lib.bfm_force_eval(C.byref(none.c_force), C.byref(pos.c_vec), C.byref(out.c_vec)) == 0
	assert (out.c_vec.data[0], out.c_vec.data[1]) == (0.0, 0.0)

	wrong = api.Vec((0, 0, 0), lib)

	assert lib.lib.bfm_force_eval(C.byref(force.c_force), C.byref(pos.c_vec), C.byref(wrong.c_vec)) == -1
	assert lib.lib.bfm_force_set_linear(C.byref(none.c_force), C.byref(wrong.c_vec)) == -1


def test_sim_run_of_kind_none_is_a_noop(lib):
	sim = api.Sim(api.CSim.NONE, lib)

	assert lib.lib.bfm_sim_run(C.byref(sim.c_sim)) == 0


def test_hot_path_fails_loudly_without_a_device(lib):
	"""no CPU fallback: without a CUDA device bfm_sim_run returns -1 and says why"""

	if ext.device_available(lib):
		pytest.skip("a CUDA device is present")

	case = cases.build("plate_40x10", lib)

	os.environ["BFM_QUIET"] = "1"

	try:
		assert lib.lib.bfm_sim_run(C.byref(case.sim.c_sim)) == -1
		assert lib.state.err.has and b"no usable CUDA device" in lib.state.err.msg
		assert not case.instance.effects.any()  # nothing was computed anywhere
	finally:
		del os.environ["BFM_QUIET"]


def test_bench_reference_arm_prints_the_contract_line():
	"""bench.py --impl reference: the reference's own CPU implementation (oracle/_ref) on a bounded sample"""

	import json
	import subprocess
	import sys

	from oracle import ref as ref_mod

	if not ref_mod.available():
		pytest.skip("oracle/_ref/libbfm_ref.so not available")

	proc = subprocess.run(
		[sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--reference-sample", "40x10"],
		capture_output=True, text=True, timeout=300,
	)

	assert proc.returncode == 0, proc.stderr[-2000:]

	lines = [line for line in proc.stdout.splitlines() if line.strip()]
	assert len(lines) == 1

	line = json.loads(lines[0])

	assert line["impl"] == "reference" and line["unit"] == "DOF/s" and line["value"] > 0
	assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
	assert line["e2e"] == {"value": line["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
	# the line describes the plate the reference really ran (the dense reference cannot hold the GPU arm's 50 M DOF)
	assert line["config"]["n_dofs"] == 2 * 41 * 11 and line["config"]["cells"] == "40x10" and line["same_config"] is False
	assert "50025002 DOF" in line["config"]["gpu_arm_workload"] and line["higher_is_better"] is True


def test_edge_derivation_parallel_path_matches_live_reference(lib, ref, tmp_path):
	"""above 2^18 half-edges the stable merge sort of bfmx_mesh_compute_edges runs as OpenMP tasks: the edge list
	(order, node pairs, element pairs) must still be the reference reader's"""

	nx, ny = 320, 150  # 96 000 triangles -> 288 000 half-edges
	coords, elems = cases.plate_arrays(nx, ny)
	path = tmp_path / "plate.obj"

	with open(path, "w") as f:
		np.savetxt(f, np.c_[coords, np.zeros(len(coords))], fmt="v %.17g %.17g %g")
		np.savetxt(f, elems.astype(np.int64) + 1, fmt="f %d %d %d")

	want = api.Mesh_wavefront(str(path), binding=ref).edges_array
	got_reader = api.Mesh_wavefront(str(path), binding=lib).edges_array
	got_generator = ext.plate(nx, ny, binding=lib, with_edges=True).edges_array

	assert len(want) > (1 << 17)
	assert np.array_equal(got_reader, want)
	assert np.array_equal(got_generator, want)


def test_internal_numbering_is_a_local_permutation(lib, monkeypatch):
	"""renumber.c: a mesh numbered at random gets a Morton numbering (a permutation; elements keep their node sets;
	the mean node-number span of an element collapses); small or well numbered meshes are left alone"""

	nx, ny = 240, 60
	coords, elems = cases.plate_arrays(nx, ny)
	rng = np.random.RandomState(3)
	perm = rng.permutation(len(coords))
	shuffled_coords = np.empty_like(coords)
	shuffled_coords[perm] = coords
	shuffled_elems = perm[elems.astype(np.int64)].astype(np.uint64)

	def span(e):
		e = np.asarray(e, np.int64)
		return float((e.max(axis=1) - e.min(axis=1)).mean())

	monkeypatch.delenv("BFM_RENUMBER", raising=False)
	assert ext.internal_numbering(api.Mesh.from_arrays(shuffled_coords, shuffled_elems, binding=lib)) is None  # 14 701 nodes: fits any cache

	monkeypatch.setenv("BFM_RENUMBER", "0")
	assert ext.internal_numbering(api.Mesh.from_arrays(shuffled_coords, shuffled_elems, binding=lib)) is None

	monkeypatch.setenv("BFM_RENUMBER", "1")
	mesh = api.Mesh.from_arrays(shuffled_coords, shuffled_elems, binding=lib)
	to_new = ext.internal_numbering(mesh)

	assert to_new is not None and np.array_equal(np.sort(to_new), np.arange(len(coords)))
	assert span(shuffled_elems) > len(coords) / 4            # random: an element spans a third of the numbering
	assert span(to_new[shuffled_elems.astype(np.int64)]) < 4 * (nx + 2)  # Morton: comparable with the row-by-row numbering's nx + 2
	assert np.array_equal(ext.internal_numbering(mesh), to_new)          # cached, deterministic


def _write_obj(path, coords, elems, odd=False, triplets=False):
	with open(path, "w") as f:
		f.write("# a comment that mentions v 1 2 3\no plate\n\n")
		np.savetxt(f, np.c_[coords, np.zeros(len(coords))], fmt="v %.17g %.17g %g")
		f.write("vn 0 0 1\nvt 0.5 0.5\ns off\n")
		half = len(elems) // 2
		np.savetxt(f, elems[:half].astype(np.int64) + 1, fmt="f %d %d %d")

		if triplets:  # "a/at/an": more than the reference's "%zu %zu %zu" reads, tolerated by both of our scanners
			np.savetxt(f, np.repeat(elems[half:].astype(np.int64) + 1, 3, axis=1), fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")
		else:
			np.savetxt(f, elems[half:].astype(np.int64) + 1, fmt="f %d %d %d")

		if odd:  # a record broken over two lines: legal for fscanf, not "one record per line"
			f.write("v 9.5 8.5\n7.5\nf 1 2\n3\n")


@pytest.mark.parametrize("odd", [False, True])
def test_parallel_wavefront_reader_matches_reference(odd, lib, ref, tmp_path, monkeypatch):
	"""files above 1 MB are parsed by all threads, one line per record, with a fallback to the serial scanner when a
	line is not exactly one record: nodes, elements and edges must be the reference reader's either way"""

	coords, elems = cases.plate_arrays(400, 160)
	path = tmp_path / "plate.obj"
	_write_obj(path, coords, elems, odd=odd)

	assert path.stat().st_size > (1 << 20)

	want = api.Mesh_wavefront(str(path), binding=ref)

	monkeypatch.delenv("BFM_READER", raising=False)
	fast = api.Mesh_wavefront(str(path), binding=lib)

	monkeypatch.setenv("BFM_READER", "serial")
	slow = api.Mesh_wavefront(str(path), binding=lib)

	assert want.n_nodes == len(coords) + odd and want.elems_array.shape[0] == len(elems) + odd

	for got in (fast, slow):
		assert np.array_equal(got.coords_array, want.coords_array)
		assert np.array_equal(got.elems_array, want.elems_array)
		assert np.array_equal(got.edges_array, want.edges_array)

	# vertex/texture/normal triplets: beyond the reference's reader; the two scanners of ours must agree
	_write_obj(path, coords, elems, odd=odd, triplets=True)

	slow = api.Mesh_wavefront(str(path), binding=lib)
	monkeypatch.delenv("BFM_READER", raising=False)
	fast = api.Mesh_wavefront(str(path), binding=lib)

	assert np.array_equal(fast.elems_array, want.elems_array) and np.array_equal(slow.elems_array, want.elems_array)
	assert np.array_equal(fast.edges_array, want.edges_array)


@pytest.mark.parametrize("kind,odd", [(3, False), (4, False), (3, True)])
def test_parallel_lepl1110_reader_matches_reference(kind, odd, lib, ref, tmp_path, monkeypatch):
	nx, ny = 300, 150
	coords, elems = cases.plate_arrays(nx, ny, kind=kind)
	boundary = np.arange(nx)  # a made-up boundary edge list: (element, node, node)
	path = tmp_path / "plate.lepl1110"

	with open(path, "w") as f:
		f.write(f"Number of nodes {len(coords)} \n")

		for i, (x, y) in enumerate(coords):
			sep = "\n   " if odd and i == 1000 else " "  # one record broken over two lines
			f.write(f"{i:6d} : {x:14.7e}{sep}{y:14.7e} \n")

		f.write(f"Number of edges {len(boundary)} \n")

		for i in boundary:
			f.write(f"{i:6d} : {i:6d} {i + 1:6d} \n")

		f.write(f"Number of {'triangles' if kind == 3 else 'quads'} {len(elems)} \n")

		for i, e in enumerate(elems):
			f.write(f"{i:6d} : " + " ".join(f"{int(v):6d}" for v in e) + " \n")

		f.write("Number of domains 2 \n")
		f.write("  Domain :      0 \n  Name : Left side \n  Number of elements :      3\n     0      1      2 \n")
		f.write("  Domain :      1 \n  Name : Entity 1 \n  Number of elements :      2\n     5      6 \n")

	assert path.stat().st_size > (1 << 20)

	want = api.Mesh_lepl1110(str(path), binding=ref)

	monkeypatch.delenv("BFM_READER", raising=False)
	fast = api.Mesh_lepl1110(str(path), binding=lib)

	monkeypatch.setenv("BFM_READER", "serial")
	slow = api.Mesh_lepl1110(str(path), binding=lib)

	for got in (fast, slow):
		assert got.kind == kind and got.n_nodes == len(coords)
		assert np.array_equal(got.coords_array, want.coords_array)
		assert np.array_equal(got.elems_array, want.elems_array)
		assert np.array_equal(got.edges_array, want.edges_array)
		assert got.c_mesh.n_domains == 2
