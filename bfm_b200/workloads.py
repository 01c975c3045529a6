"""The synthetic workload BASELINE.json's metric is quoted on: the structured triangulated plate of
SURVEY.md section 8(d), built through the pybfm-style object model on any binding (ours or the
compiled reference - same C ABI).

    rectangle [0,4]x[0,1], (nx+1)(ny+1) nodes row-major, every cell split into triangles (a,b,d),(a,d,c);
    steel E=211e9, nu=0.3, rho=7850; plane stress; gravity (0,-9.81); left edge clamped in X and Y.
    Deterministic: no random numbers, hence no seed.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import api

STEEL = (211.0e9, 0.3, 7.85e3)
GRAVITY = (0.0, -9.81)


@dataclass
class PlateCase:
	nx: int
	ny: int
	mesh: api.Mesh
	sim: api.CSim
	instance: api.CInstance
	keep: list = field(default_factory=list)

	@property
	def n_dofs(self) -> int:
		return 2 * self.mesh.n_nodes


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


def plate_case(nx: int, ny: int, kind: int = 3, binding: api.Binding | None = None, native_mesh: bool = True) -> PlateCase:
	"""the cantilever plate as a ready-to-run Sim.  native_mesh=True generates the mesh inside our
	library (bfmx_mesh_plate, multi-threaded - what the large sizes need); False goes through numpy
	arrays and works on the reference binding too."""

	binding = binding if binding is not None else api.default_binding()

	if native_mesh:
		from . import ext

		mesh = ext.plate(nx, ny, kind=kind, binding=binding)
	else:
		coords, elems = plate_arrays(nx, ny, kind=kind)
		mesh = api.Mesh.from_arrays(coords, elems, binding=binding)

	# left edge i = 0: nodes j * (nx + 1)
	left = np.zeros(mesh.n_nodes, dtype=np.bool_)
	left[:: nx + 1] = True

	E, nu, rho = STEEL
	material = api.Material("steel", rho, E, nu, binding=binding)
	rule = api.Rule_gauss_legendre(2, mesh.kind, binding=binding)
	obj = api.Obj(mesh, material, rule)
	instance = api.Instance(obj)

	for cond_kind in (api.Condition.DIRICHLET_X, api.Condition.DIRICHLET_Y):
		cond = api.Condition(mesh, cond_kind, 0.0)
		cond.set_nodes(left)
		instance.add_condition(cond)

	sim = api.Sim(api.CSim.PLANAR_STRESS, binding=binding)
	sim.add_instance(instance)
	force = api.Force_linear(GRAVITY, binding=binding)
	sim.add_force(force)

	return PlateCase(nx, ny, mesh, sim, instance, [material, rule, obj, force])


def effects_view(instance: api.CInstance) -> np.ndarray:
	"""instance->effects without a copy (read right after bfm_sim_run)"""

	n = instance.c_instance.n_effects
	return np.ctypeslib.as_array(instance.c_instance.effects, shape=(n,))


@dataclass
class SweepCase:
	sim: api.CSim
	instance: api.CInstance
	keep: list = field(default_factory=list)


def lepl_sweep(mesh_path: str, problem_path: str, count: int, binding: api.Binding | None = None) -> list[SweepCase]:
	"""`count` variants of one LEPL1110 problem on one shared mesh, Young's modulus scaled by 1 + i / count
	(a design sweep: BASELINE.json configs[4]).  Conditions and gravity are those the problem file gives."""

	binding = binding if binding is not None else api.default_binding()
	mesh = api.Mesh_lepl1110(mesh_path, binding=binding)
	ez = api.Ez_lepl1110(mesh, problem_path)
	m = ez.c_ez.material
	conds = ez.conditions()
	g = ez.c_ez.gravity.linear.force
	gravity = (g.data[0], g.data[1]) if ez.c_ez.sim.n_forces else None
	out = []

	for i in range(count):
		material = api.Material(f"sweep{i}", m.rho, m.E * (1.0 + i / count), m.nu, binding=binding)
		rule = api.Rule_gauss_legendre(2, mesh.kind, binding=binding)
		obj = api.Obj(mesh, material, rule)
		instance = api.Instance(obj)
		keep = [mesh, ez, material, rule, obj]

		for kind, value, mask in conds:
			cond = api.Condition(mesh, kind, value)
			cond.set_nodes(mask.astype(bool))
			instance.add_condition(cond)

		sim = api.Sim(ez.sim.kind, binding=binding)
		sim.add_instance(instance)

		if gravity is not None:
			force = api.Force_linear(gravity, binding=binding)
			sim.add_force(force)
			keep.append(force)

		out.append(SweepCase(sim, instance, keep))

	return out
