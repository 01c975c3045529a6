"""The parity cases: BASELINE.json's configs at oracle-feasible sizes, plus edge cases.

Every case is built through the pybfm-style object model (bfm_b200.api) on a given binding, so the same
recipe drives our library, the compiled reference (oracle/_ref) and - via ``oracle_problem`` - the C
restatement.  ``tests/golden/make_golden.py`` runs all of them on the reference and stores the results.
"""

from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from bfm_b200 import api

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> does make_golden also store the dense-derived data (pattern digest, nnz, ...) ?
CASES = {
	"lepl8": True,              # config 1: lepl1110.py meshes/8.lepl1110 problems/problem.txt
	"lepl8_all_kinds": True,    # all 8 condition kinds, non-zero Dirichlet values, interleaved order
	"lepl8_axisym": True,       # axisymmetric path
	"bridge": True,             # config 2: bridge.obj, plane stress, AA7075, gravity, 16 clamped nodes
	"bridge_dam": True,         # the mesh examples/deformation.py actually loads, 64 clamped nodes
	"plate_40x10": True,        # config 4 at oracle sizes
	"plate_80x20": True,
	"plate_160x40": True,       # SURVEY.md 8(d) parity size: 13 202 DOF (the reference arm's sample in bench.py)
	"plate_300x75": False,      # 45 752 DOF, just under the dense reference's ceiling of 46 340
	"plate_q4_24x6": True,      # quads
	"plate_funky_20x5": True,   # FUNKY + LINEAR forces together (order of accumulation)
	"plate_neumann_16x4": True, # Neumann tip load on generated edges
	"gear60": False,            # config 3: irregular sparsity, mixed Dirichlet/Neumann (16 410 DOF)
	"plate_jitter_40x10": True,     # interior nodes moved by +-0.2 h (seed 0): irregular geometry, no exact cancellations
	"plate_q4_jitter_24x6": True,   # the same on quads: the Jacobian differs at every Gauss point; plane strain
}

HEAVY = {"gear60", "plate_300x75"}  # seconds-to-minutes on the dense reference


@dataclass
class Case:
	name: str
	mesh: api.Mesh
	sim: api.CSim
	instance: api.CInstance
	sim_kind: int
	material: tuple            # (E, nu, rho)
	forces: list               # constants (fx, fy) or callables f(x, y)
	conditions: list           # (kind, value, mask)
	keep: list = field(default_factory=list)


def plate_arrays(nx: int, ny: int, lx: float = 4.0, ly: float = 1.0, kind: int = 3):
	"""numpy twin of bfmx_mesh_plate (bfm_b200/csrc/mesh.c): identical coordinates and connectivity"""

	i = np.arange(nx + 1, dtype=np.float64)
	j = np.arange(ny + 1, dtype=np.float64)
	coords = np.zeros(((ny + 1) * (nx + 1), 2))
	coords[:, 0] = np.tile(lx * i / nx, ny + 1)
	coords[:, 1] = np.repeat(ly * j / ny, nx + 1)

	a = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
	b, c, d = a + 1, a + nx + 1, a + nx + 2

	if kind == 3:
		elems = np.stack([a, b, d, a, d, c], axis=1).reshape(-1, 3)
	else:
		elems = np.stack([d, c, a, b], axis=1)

	return coords, elems.astype(np.uint64)


def jittered_plate_arrays(nx: int, ny: int, kind: int = 3, amplitude: float = 0.2, seed: int = 0):
	"""the structured plate with every interior node moved by a uniform +-amplitude * h in x and y
	(numpy RandomState(seed): the same numbers wherever the fixtures are regenerated)"""

	coords, elems = plate_arrays(nx, ny, kind=kind)
	hx, hy = 4.0 / nx, 1.0 / ny
	rng = np.random.RandomState(seed)
	shift = rng.uniform(-amplitude, amplitude, size=coords.shape) * np.array([hx, hy])

	i = np.tile(np.arange(nx + 1), ny + 1)
	j = np.repeat(np.arange(ny + 1), nx + 1)
	interior = (i > 0) & (i < nx) & (j > 0) & (j < ny)

	coords = coords.copy()
	coords[interior] += shift[interior]

	return coords, elems


def _plate_mesh(binding, nx, ny, kind=3, edges=False):
	coords, elems = plate_arrays(nx, ny, kind=kind)
	mesh = api.Mesh.from_arrays(coords, elems, binding=binding)

	if edges:
		mesh_edges = plate_edges(mesh)
		mesh = api.Mesh.from_arrays(coords, elems, edges=mesh_edges, binding=binding)

	return mesh


def plate_edges(mesh) -> np.ndarray:
	"""edges of an in-memory mesh in the reference's own order: write a temporary OBJ, read it back"""

	import tempfile

	with tempfile.NamedTemporaryFile("w", suffix=".obj", delete=False) as f:
		for x, y in mesh.coords_array:
			f.write(f"v {x!r} {y!r} 0\n")

		for tri in mesh.elems_array:
			f.write("f " + " ".join(str(int(v) + 1) for v in tri) + "\n")

		path = f.name

	try:
		return api.Mesh_wavefront(path, binding=mesh.binding).edges_array
	finally:
		os.unlink(path)


def _assemble_case(name, binding, mesh, sim_kind, material, forces, conditions) -> Case:
	E, nu, rho = material
	mat = api.Material("case", rho, E, nu, binding=binding)
	rule = api.Rule_gauss_legendre(2, mesh.kind, binding=binding)
	obj = api.Obj(mesh, mat, rule)
	instance = api.Instance(obj)
	keep = [mat, rule, obj]

	for kind, value, mask in conditions:
		cond = api.Condition(mesh, kind, value)
		cond.set_nodes(mask)
		instance.add_condition(cond)

	sim = api.Sim(sim_kind, binding=binding)
	sim.add_instance(instance)

	for f in forces:
		force = api.Force_funky(f, binding=binding) if callable(f) else api.Force_linear(f, binding=binding)
		sim.add_force(force)

	return Case(name, mesh, sim, instance, sim_kind, material, list(forces), list(conditions), keep)


def _ez_case(name, binding, mesh_file, problem_file) -> Case:
	mesh = api.Mesh_lepl1110(os.path.join(GOLDEN, "meshes", mesh_file), binding=binding)
	ez = api.Ez_lepl1110(mesh, os.path.join(GOLDEN, "problems", problem_file))

	g = ez.c_ez.gravity.linear.force
	forces = [(g.data[0], g.data[1])] if ez.c_ez.sim.n_forces else []
	m = ez.c_ez.material

	return Case(name, mesh, ez.sim, ez.instance, ez.sim.kind, (m.E, m.nu, m.rho), forces, ez.conditions(), [ez])


STEEL = (211.0e9, 0.3, 7.85e3)
AA7075 = (71.7e9, 0.33, 2.81e3)
GRAVITY = (0.0, -9.81)


def _lowest_nodes(mesh, count):
	y = mesh.coords_array[:, 1]
	order = np.lexsort((np.arange(len(y)), y))
	mask = np.zeros(len(y), bool)
	mask[order[:count]] = True
	return mask


def build(name: str, binding=None) -> Case:
	binding = binding if binding is not None else api.default_binding()

	if name == "lepl8":
		return _ez_case(name, binding, "8.lepl1110", "problem.txt")

	if name == "lepl8_all_kinds":
		return _ez_case(name, binding, "8.lepl1110", "lepl8_all_kinds.txt")

	if name == "lepl8_axisym":
		return _ez_case(name, binding, "8.lepl1110", "lepl8_axisym.txt")

	if name == "gear60":
		return _ez_case(name, binding, "gear60_full.lepl1110", "gear60_mixed.txt")

	if name in ("bridge", "bridge_dam"):
		mesh = api.Mesh_wavefront(os.path.join(GOLDEN, "meshes", "bridge.obj" if name == "bridge" else "bridge-dam.obj"), binding=binding)
		mask = _lowest_nodes(mesh, 16 if name == "bridge" else 64)
		conds = [(api.Condition.DIRICHLET_X, 0.0, mask), (api.Condition.DIRICHLET_Y, 0.0, mask)]
		return _assemble_case(name, binding, mesh, api.CSim.PLANAR_STRESS, AA7075, [GRAVITY], conds)

	if name.startswith("plate_") and "jitter" in name:
		kind = 4 if "_q4_" in name else 3
		nx, ny = (int(v) for v in name.rsplit("_", 1)[1].split("x"))
		coords, elems = jittered_plate_arrays(nx, ny, kind)
		mesh = api.Mesh.from_arrays(coords, elems, binding=binding)
		left = coords[:, 0] == 0.0
		conds = [(api.Condition.DIRICHLET_X, 0.0, left), (api.Condition.DIRICHLET_Y, 0.0, left)]
		sim_kind = api.CSim.PLANAR_STRAIN if kind == 4 else api.CSim.PLANAR_STRESS

		return _assemble_case(name, binding, mesh, sim_kind, STEEL, [GRAVITY], conds)

	if name.startswith("plate_"):
		kind = 4 if "_q4_" in name else 3
		nx, ny = (int(v) for v in name.rsplit("_", 1)[1].split("x"))
		neumann = "neumann" in name
		mesh = _plate_mesh(binding, nx, ny, kind, edges=neumann)

		x = mesh.coords_array[:, 0]
		left = x == 0.0
		conds = [(api.Condition.DIRICHLET_X, 0.0, left), (api.Condition.DIRICHLET_Y, 0.0, left)]
		forces = [GRAVITY]

		if "funky" in name:
			forces = [lambda px, py: (100.0 * py, -9.81 * (1.0 + px)), GRAVITY, lambda px, py: (px * py, 0.25)]

		if neumann:
			right = x == x.max()
			conds.append((api.Condition.NEUMANN_Y, -2.0e6, right))
			conds.append((api.Condition.NEUMANN_NORMAL, 1.0e5, right))

		return _assemble_case(name, binding, mesh, api.CSim.PLANAR_STRESS, STEEL, forces, conds)

	raise KeyError(name)


def oracle_problem(case: Case):
	"""the same case for the C restatement (oracle/bfm_oracle.c)"""

	from oracle import orc

	coords = case.mesh.coords_array
	forces = []

	for f in case.forces:
		if callable(f):
			forces.append(np.array([f(px, py) for px, py in coords], dtype=np.float64))
		else:
			forces.append(tuple(f))

	E, nu, rho = case.material

	return orc.Problem(coords, case.mesh.elems_array, case.sim_kind, E, nu, rho, forces=forces, conditions=case.conditions, edges=case.mesh.edges_array)


def golden() -> dict:
	return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


# ---- the same cases without any libbfm: straight from the fixture files into the oracle ------------


def _read_lepl(path):
	"""minimal LEPL1110 reader (python) for the oracle-only path"""

	tok = open(path).read().split("\n")
	pos = 0

	def section(prefix):
		nonlocal pos
		assert tok[pos].startswith(prefix), (tok[pos], prefix)
		count = int(tok[pos].split()[-1])
		rows = [line.split(":")[1].split() for line in tok[pos + 1:pos + 1 + count]]
		pos += 1 + count
		return rows

	coords = np.array(section("Number of nodes"), dtype=np.float64)
	edge_rows = [(line.split(":")[0], line.split(":")[1].split()) for line in tok[pos + 1:pos + 1 + int(tok[pos].split()[-1])]]
	edges = np.array([[int(n[0]), int(n[1]), int(e), -1] for e, n in edge_rows], dtype=np.int64)
	pos += 1 + len(edge_rows)
	elems = np.array(section("Number of"), dtype=np.uint64)

	domains = {}
	n_domains = int(tok[pos].split()[-1])
	pos += 1

	for _ in range(n_domains):
		name = tok[pos + 1].split(":", 1)[1].lstrip(" ")
		count = int(tok[pos + 2].split(":")[1])
		pos += 3
		ids = []

		while len(ids) < count:
			ids += [int(v) for v in tok[pos].split()]
			pos += 1

		domains[name] = ids

	return coords, elems, edges, domains


_KINDS = {
	"dirichlet-x": 0, "dirichlet-y": 1, "neumann-x": 2, "neumann-y": 3,
	"neumann-normal": 4, "neumann-tangent": 5, "dirichlet-normal": 6, "dirichlet-tangent": 7,
}


def _read_problem(path, n_nodes, edges, domains):
	sim_kind, E, nu, rho, g, conds = 0, 0.0, 0.0, 0.0, None, []

	for line in open(path):
		key, _, value = line.partition(":")
		key = key.strip().lower()
		value = value.strip("\n").lstrip(" ")

		if key == "type of problem":
			sim_kind = {"planar strain": 1, "planar stress": 2, "axi-symetric ": 3}[value[:13].lower()]
		elif key == "young modulus":
			E = float(value)
		elif key == "poisson ratio":
			nu = float(value)
		elif key == "mass density":
			rho = float(value)
		elif key == "gravity":
			g = -float(value)
		elif key == "boundary condition":
			kind, rest = value.split("=")
			val, dom = rest.split(":", 1)
			dom = dom.lstrip(" ")
			mask = np.zeros(n_nodes, np.uint8)

			for name, ids in domains.items():
				if name[:25].lower() == dom[:25].lower():
					for e in ids:
						mask[edges[e, 0]] = mask[edges[e, 1]] = 1

					break

			conds.append((_KINDS[kind.strip().lower()], float(val), mask))

	return sim_kind, E, nu, rho, g, conds


def build_oracle_only(name: str):
	"""orc.Problem for a case using no libbfm at all (neither ours nor the reference)"""

	from oracle import orc

	ez = {
		"lepl8": ("8.lepl1110", "problem.txt"),
		"lepl8_all_kinds": ("8.lepl1110", "lepl8_all_kinds.txt"),
		"lepl8_axisym": ("8.lepl1110", "lepl8_axisym.txt"),
		"gear60": ("gear60_full.lepl1110", "gear60_mixed.txt"),
	}

	if name in ez:
		coords, elems, edges, domains = _read_lepl(os.path.join(GOLDEN, "meshes", ez[name][0]))
		sim_kind, E, nu, rho, g, conds = _read_problem(os.path.join(GOLDEN, "problems", ez[name][1]), len(coords), edges, domains)
		return orc.Problem(coords, elems, sim_kind, E, nu, rho, forces=[(0.0, g)] if g is not None else [], conditions=conds, edges=edges)

	if name.startswith("plate_") and "jitter" in name:
		kind = 4 if "_q4_" in name else 3
		nx, ny = (int(v) for v in name.rsplit("_", 1)[1].split("x"))
		coords, elems = jittered_plate_arrays(nx, ny, kind)
		left = (coords[:, 0] == 0.0).astype(np.uint8)
		E, nu, rho = STEEL

		return orc.Problem(coords, elems, 1 if kind == 4 else 2, E, nu, rho, forces=[GRAVITY], conditions=[(0, 0.0, left), (1, 0.0, left)])

	if name.startswith("plate_") and "neumann" not in name:
		kind = 4 if "_q4_" in name else 3
		nx, ny = (int(v) for v in name.rsplit("_", 1)[1].split("x"))
		coords, elems = plate_arrays(nx, ny, kind=kind)
		left = (coords[:, 0] == 0.0).astype(np.uint8)
		conds = [(0, 0.0, left), (1, 0.0, left)]
		forces = [GRAVITY]

		if "funky" in name:
			fns = [lambda px, py: (100.0 * py, -9.81 * (1.0 + px)), None, lambda px, py: (px * py, 0.25)]
			forces = [GRAVITY if f is None else np.array([f(px, py) for px, py in coords]) for f in fns]

		E, nu, rho = STEEL
		return orc.Problem(coords, elems, 2, E, nu, rho, forces=forces, conditions=conds)

	# cases that need a libbfm reader for their inputs (OBJ edges): use ours
	return oracle_problem(build(name))
