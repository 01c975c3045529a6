"""Host-side mirror of pybfm's object model, bound through ctypes.

Same class names, constructor arguments and error behaviour (``assert not lib.bfm_*(...)``) as the
reference wrapper (``pybfm/bfm/{mesh,condition,force,material,rule,obj,instance,sim,ez,vec}.py``),
minus the OpenGL viewer.  Every class takes an optional ``binding`` so that the very same driver
code can run on the reference library (``oracle/_ref/libbfm_ref.so``) in the parity tests.
"""

from __future__ import annotations

import ctypes as C
import os
from collections.abc import Callable

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbfm.so")

_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


class Binding:
	"""One loaded libbfm + the default state all objects share (pybfm/bfm/state.py:3-4)."""

	def __init__(self, path: str, extra: dict | None = None):
		self.path = path
		self.lib = abi.bind(path, extra)
		self.state = abi.State()
		self.lib.bfm_state_create(C.byref(self.state))


_default: Binding | None = None


def default_binding() -> Binding:
	"""The product library.  Fails loudly when it has not been built - there is no fallback."""

	global _default

	if _default is None:
		if not os.path.exists(LIB_PATH):
			raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")

		from . import ext

		_default = Binding(LIB_PATH, ext.PROTOTYPES)

	return _default


def _b(binding: Binding | None) -> Binding:
	return binding if binding is not None else default_binding()


def _malloc_copy(arr: np.ndarray) -> int:
	"""libc-malloc'd copy of arr (the mesh destructor releases buffers with state->free)."""

	arr = np.ascontiguousarray(arr)
	ptr = _libc.malloc(max(arr.nbytes, 1))
	assert ptr
	C.memmove(ptr, arr.ctypes.data, arr.nbytes)
	return ptr


# ---------------------------------------------------------------------------------------------
# meshes (pybfm/bfm/mesh.py)
# ---------------------------------------------------------------------------------------------


class Mesh:
	SIMPLEX = 3
	QUAD = 4

	def __init__(self, dim: int, kind: int, binding: Binding | None = None):
		self.binding = _b(binding)
		self.c_mesh = abi.Mesh()
		assert not self.binding.lib.bfm_mesh_create(C.byref(self.c_mesh), C.byref(self.binding.state), dim, kind)

		self.dim = dim
		self.kind = kind

	def __del__(self):
		try:
			self.binding.lib.bfm_mesh_destroy(C.byref(self.c_mesh))
		except Exception:
			pass

	@classmethod
	def from_arrays(cls, coords, elems, edges=None, binding: Binding | None = None) -> "Mesh":
		"""In-memory mesh: coords [n_nodes, 2] float64, elems [n_elems, kind] ints,
		edges optional [n_edges, 4] = (node0, node1, elem0, elem1)."""

		coords = np.ascontiguousarray(coords, dtype=np.float64)
		elems = np.ascontiguousarray(elems, dtype=np.uint64)
		mesh = cls(coords.shape[1], elems.shape[1], binding)
		m = mesh.c_mesh

		m.n_nodes = coords.shape[0]
		m.n_elems = elems.shape[0]
		m.coords = C.cast(_malloc_copy(coords), abi.c_double_p)
		m.elems = C.cast(_malloc_copy(elems), abi.c_size_t_p)

		if edges is not None:
			edges = np.ascontiguousarray(edges, dtype=np.int64)
			m.n_edges = edges.shape[0]
			m.edges = C.cast(_malloc_copy(edges), C.POINTER(abi.Edge))

		return mesh

	# numpy views of the C buffers

	@property
	def n_nodes(self) -> int:
		return self.c_mesh.n_nodes

	@property
	def n_elems(self) -> int:
		return self.c_mesh.n_elems

	@property
	def coords_array(self) -> np.ndarray:
		m = self.c_mesh
		return np.ctypeslib.as_array(m.coords, shape=(m.n_nodes, m.dim)).copy() if m.n_nodes else np.zeros((0, m.dim))

	@property
	def elems_array(self) -> np.ndarray:
		m = self.c_mesh
		return np.ctypeslib.as_array(m.elems, shape=(m.n_elems, m.kind)).copy() if m.n_elems else np.zeros((0, m.kind), np.uint64)

	@property
	def edges_array(self) -> np.ndarray:
		m = self.c_mesh

		if not m.n_edges:
			return np.zeros((0, 4), np.int64)

		raw = np.ctypeslib.as_array(C.cast(m.edges, C.POINTER(C.c_int64)), shape=(m.n_edges, 4))
		return raw.copy()

	@property
	def coords(self):  # pybfm/bfm/mesh.py:25-27
		return self.coords_array.reshape(-1).tolist()

	def domains(self) -> dict[str, np.ndarray]:
		m = self.c_mesh
		out = {}

		for i in range(m.n_domains):
			d = m.domains[i]
			out[d.name.decode()] = np.ctypeslib.as_array(d.elements, shape=(d.n_elements,)).copy() if d.n_elements else np.zeros(0, np.uint64)

		return out


class Mesh_lepl1110(Mesh):
	def __init__(self, name: str, binding: Binding | None = None):
		self.binding = _b(binding)
		self.c_mesh = abi.Mesh()
		assert not self.binding.lib.bfm_mesh_read_lepl1110(C.byref(self.c_mesh), C.byref(self.binding.state), name.encode())

		self.dim = self.c_mesh.dim
		self.kind = self.c_mesh.kind


class Mesh_wavefront(Mesh):
	def __init__(self, name: str, full: bool = False, binding: Binding | None = None):
		self.binding = _b(binding)
		self.c_mesh = abi.Mesh()
		assert not self.binding.lib.bfm_mesh_read_wavefront(C.byref(self.c_mesh), C.byref(self.binding.state), name.encode(), full)

		self.dim = self.c_mesh.dim
		self.kind = self.c_mesh.kind


# ---------------------------------------------------------------------------------------------
# conditions, vectors, forces (pybfm/bfm/{condition,vec,force}.py)
# ---------------------------------------------------------------------------------------------


class Condition:
	DIRICHLET_X = 0
	DIRICHLET_Y = 1
	NEUMANN_X = 2
	NEUMANN_Y = 3
	NEUMANN_NORMAL = 4
	NEUMANN_TANGENT = 5
	DIRICHLET_NORMAL = 6
	DIRICHLET_TANGENT = 7

	def __init__(self, mesh: Mesh, kind: int, value: float = 0.0):
		self.binding = mesh.binding
		self.c_condition = abi.Condition()
		assert not self.binding.lib.bfm_condition_create(C.byref(self.c_condition), C.byref(self.binding.state), C.byref(mesh.c_mesh), kind)

		self.c_condition.value = value
		self.mesh = mesh

	def __del__(self):
		try:
			self.binding.lib.bfm_condition_destroy(C.byref(self.c_condition))
		except Exception:
			pass

	def populate(self, discriminator: Callable):  # pybfm/bfm/condition.py:26-29
		coords = self.mesh.coords_array

		for i in range(self.mesh.c_mesh.n_nodes):
			self.c_condition.nodes[i] = bool(discriminator(self.mesh, tuple(coords[i])))

	def set_nodes(self, mask):
		mask = np.ascontiguousarray(mask, dtype=np.bool_)
		assert mask.shape == (self.mesh.c_mesh.n_nodes,)
		C.memmove(self.c_condition.nodes, mask.ctypes.data, mask.nbytes)

	@property
	def nodes_array(self) -> np.ndarray:
		n = self.mesh.c_mesh.n_nodes
		return np.ctypeslib.as_array(C.cast(self.c_condition.nodes, C.POINTER(C.c_uint8)), shape=(n,)).copy()


class Vec:
	def __init__(self, vec, binding: Binding | None = None):
		self.binding = _b(binding)
		self.c_vec = abi.Vec()
		assert not self.binding.lib.bfm_vec_create(C.byref(self.c_vec), C.byref(self.binding.state), len(vec))

		for i, val in enumerate(vec):
			self.c_vec.data[i] = val

	def __del__(self):
		try:
			self.binding.lib.bfm_vec_destroy(C.byref(self.c_vec))
		except Exception:
			pass


class Force:
	def __init__(self, dim: int, binding: Binding | None = None):
		self.binding = _b(binding)
		self.c_force = abi.Force()
		assert not self.binding.lib.bfm_force_create(C.byref(self.c_force), C.byref(self.binding.state), dim)

	def __del__(self):
		try:
			self.binding.lib.bfm_force_destroy(C.byref(self.c_force))
		except Exception:
			pass


class Force_none(Force):
	def __init__(self, dim: int, binding: Binding | None = None):
		super().__init__(dim, binding)
		assert not self.binding.lib.bfm_force_set_none(C.byref(self.c_force))


class Force_linear(Force):
	def __init__(self, vec, binding: Binding | None = None):
		super().__init__(len(vec), binding)

		c_vector = Vec(vec, self.binding)
		assert not self.binding.lib.bfm_force_set_linear(C.byref(self.c_force), C.byref(c_vector.c_vec))

	@classmethod
	def earth_gravity_2d(cls, binding: Binding | None = None):  # Force_linear.EARTH_GRAVITY_2D, pybfm/bfm/force.py:33
		return cls((0, -9.81), binding)


class Force_funky(Force):
	"""Position-dependent force from a Python callable ``f(x, y) -> (fx, fy)`` (the reference
	declares the C hook, bfm/force.h:18, but never bound it: pybfm/bfm/force.py:37)."""

	def __init__(self, func: Callable, dim: int = 2, binding: Binding | None = None):
		super().__init__(dim, binding)

		def trampoline(_force, pos, ref, _data):
			out = func(*[pos.contents.data[i] for i in range(dim)])

			for i in range(dim):
				ref.contents.data[i] = out[i]

			return 0

		self._cb = abi.FORCE_FUNKY_FN(trampoline)
		assert not self.binding.lib.bfm_force_set_funky(C.byref(self.c_force), self._cb, None)


# ---------------------------------------------------------------------------------------------
# material, rule, obj, instance, sim (pybfm/bfm/{material,rule,obj,instance,sim}.py)
# ---------------------------------------------------------------------------------------------


class CMaterial:
	def __init__(self, c_material, binding: Binding):
		self.c_material = c_material
		self.binding = binding

	@property
	def name(self):
		return C.string_at(self.c_material.name)


class Material(CMaterial):
	def __init__(self, name: str, rho: float, E: float, nu: float, colour=(1, 1, 1, 1), binding: Binding | None = None):
		binding = _b(binding)
		c_material = abi.Material()
		assert not binding.lib.bfm_material_create(C.byref(c_material), C.byref(binding.state), name.encode(), rho, E, nu)
		assert not binding.lib.bfm_material_set_colour(C.byref(c_material), *colour)

		super().__init__(c_material, binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_material_destroy(C.byref(self.c_material))
		except Exception:
			pass

	@classmethod
	def aa7075(cls, binding: Binding | None = None):  # Material.AA7075, pybfm/bfm/material.py:18
		return cls("AA7075", 2.81e3, 71.7e9, 0.33, (0.667, 0.439, 0.459, 1), binding)

	@classmethod
	def steel(cls, binding: Binding | None = None):  # the LEPL1110 course material (problems/problem.txt)
		return cls("Steel", 7.85e3, 211.0e9, 0.3, binding=binding)


class CRule:
	def __init__(self, c_rule, binding: Binding):
		self.c_rule = c_rule
		self.binding = binding


class Rule(CRule):
	def __init__(self, dim: int, kind: int, n_points: int, binding: Binding | None = None):
		binding = _b(binding)
		c_rule = abi.Rule()
		assert not binding.lib.bfm_rule_create(C.byref(c_rule), C.byref(binding.state), dim, kind, n_points)

		super().__init__(c_rule, binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_rule_destroy(C.byref(self.c_rule))
		except Exception:
			pass


class Rule_gauss_legendre(CRule):
	def __init__(self, dim: int, kind: int, binding: Binding | None = None):
		binding = _b(binding)
		c_rule = abi.Rule()
		assert not binding.lib.bfm_rule_create_gauss_legendre(C.byref(c_rule), C.byref(binding.state), dim, kind)

		super().__init__(c_rule, binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_rule_destroy(C.byref(self.c_rule))
		except Exception:
			pass

	def tables(self):
		r = self.c_rule
		w = np.array([r.weights[i] for i in range(r.n_points)])
		p = np.array([[r.points[i][k] for k in range(r.dim)] for i in range(r.n_points)])
		return w, p


class CObj:
	def __init__(self, c_obj, mesh: Mesh, material: CMaterial, rule: CRule):
		self.c_obj = c_obj

		self.mesh = mesh
		self.material = material
		self.rule = rule


class Obj(CObj):
	def __init__(self, mesh: Mesh, material: CMaterial, rule: CRule):
		binding = mesh.binding
		c_obj = abi.Obj()

		assert not binding.lib.bfm_obj_create(
			C.byref(c_obj), C.byref(binding.state), C.byref(mesh.c_mesh), C.byref(material.c_material), C.byref(rule.c_rule)
		)

		self.binding = binding
		super().__init__(c_obj, mesh, material, rule)

	def __del__(self):
		try:
			self.binding.lib.bfm_obj_destroy(C.byref(self.c_obj))
		except Exception:
			pass


class CInstance:
	def __init__(self, c_instance, obj: CObj, binding: Binding):
		self.c_instance = c_instance
		self.obj = obj
		self.binding = binding
		self._conditions: list[Condition] = []

	def add_condition(self, condition: Condition):
		assert not self.binding.lib.bfm_instance_add_condition(C.byref(self.c_instance), C.byref(condition.c_condition))
		self._conditions.append(condition)  # keep the borrowed pointer alive

	@property
	def effects(self) -> np.ndarray:
		"""Displacements as an [n_nodes, dim] array (pybfm/bfm/instance.py:85-93 returns the flat list)."""

		n = self.c_instance.n_effects
		dim = self.obj.mesh.c_mesh.dim
		return np.ctypeslib.as_array(self.c_instance.effects, shape=(n,)).copy().reshape(-1, dim)


class Instance(CInstance):
	def __init__(self, obj: Obj):
		binding = obj.mesh.binding
		c_instance = abi.Instance()
		assert not binding.lib.bfm_instance_create(C.byref(c_instance), C.byref(binding.state), C.byref(obj.c_obj))

		super().__init__(c_instance, obj, binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_instance_destroy(C.byref(self.c_instance))
		except Exception:
			pass


class CSim:
	NONE = 0
	PLANAR_STRAIN = 1
	PLANAR_STRESS = 2
	AXISYMMETRIC_STRAIN = 3

	def __init__(self, c_sim, instances: list, binding: Binding):
		self.c_sim = c_sim
		self.instances = instances
		self.binding = binding
		self._forces: list[Force] = []

	@property
	def kind(self) -> int:
		return self.c_sim.kind

	def add_instance(self, instance: CInstance):
		assert not self.binding.lib.bfm_sim_add_instance(C.byref(self.c_sim), C.byref(instance.c_instance))
		self.instances.append(instance)

	def add_force(self, force: Force):
		assert not self.binding.lib.bfm_sim_add_force(C.byref(self.c_sim), C.byref(force.c_force))
		self._forces.append(force)

	def run(self):  # pybfm/bfm/sim.py:33-34
		assert not self.binding.lib.bfm_sim_run(C.byref(self.c_sim))


class Sim(CSim):
	def __init__(self, kind: int, binding: Binding | None = None):
		binding = _b(binding)
		c_sim = abi.Sim()
		assert not binding.lib.bfm_sim_create(C.byref(c_sim), C.byref(binding.state), kind)

		super().__init__(c_sim, [], binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_sim_destroy(C.byref(self.c_sim))
		except Exception:
			pass


# ---------------------------------------------------------------------------------------------
# LEPL1110 convenience (pybfm/bfm/ez.py)
# ---------------------------------------------------------------------------------------------


class _Borrowed:
	"""wrap a struct embedded in bfm_ez_lepl1110_t without copying it"""

	@staticmethod
	def field(struct, name):
		cls = dict(struct._fields_)[name]
		return cls.from_buffer(struct, getattr(type(struct), name).offset)


class Ez_lepl1110:
	def __init__(self, mesh: Mesh, name: str):
		self.binding = mesh.binding
		self.c_ez = abi.Ez()  # ctypes zero-fills, like ffi.new

		assert not self.binding.lib.bfm_ez_lepl1110_create(C.byref(self.c_ez), C.byref(self.binding.state), C.byref(mesh.c_mesh), name.encode())

		self.mesh = mesh
		self.material = CMaterial(_Borrowed.field(self.c_ez, "material"), self.binding)
		self.rule = CRule(_Borrowed.field(self.c_ez, "rule"), self.binding)
		self.obj = CObj(_Borrowed.field(self.c_ez, "obj"), self.mesh, self.material, self.rule)
		self.instance = CInstance(_Borrowed.field(self.c_ez, "instance"), self.obj, self.binding)
		self.sim = CSim(_Borrowed.field(self.c_ez, "sim"), [self.instance], self.binding)

	def __del__(self):
		try:
			self.binding.lib.bfm_ez_lepl1110_destroy(C.byref(self.c_ez))
		except Exception:
			pass

	def write(self, filename: str, shift: int):
		assert not self.binding.lib.bfm_ez_lepl1110_write(C.byref(self.c_ez), shift, filename.encode())

	def conditions(self):
		"""(kind, value, node mask) of every parsed boundary condition"""

		out = []
		n = self.mesh.c_mesh.n_nodes

		for i in range(self.c_ez.n_conditions):
			c = self.c_ez.conditions[i]
			mask = np.ctypeslib.as_array(C.cast(c.nodes, C.POINTER(C.c_uint8)), shape=(n,)).copy()
			out.append((c.kind, c.value, mask))

		return out


# ---------------------------------------------------------------------------------------------
# staged access: bfm_system_* (bfm/system.h) - what a C caller driving the pieces would do
# ---------------------------------------------------------------------------------------------


class System:
	"""assemble (+BC) -> renumber -> solve, step by step"""

	def __init__(self, sim: CSim, instance: CInstance | None = None):
		self.binding = sim.binding
		self.sim = sim
		self.instance = instance if instance is not None else sim.instances[0]
		self.c_system = abi.System()

		lib = self.binding.lib

		create = {
			CSim.PLANAR_STRAIN: lib.bfm_system_create_planar_strain,
			CSim.PLANAR_STRESS: lib.bfm_system_create_planar_stress,
			CSim.AXISYMMETRIC_STRAIN: lib.bfm_system_create_axisymmetric_strain,
		}[sim.kind]

		assert not create(C.byref(self.c_system), C.byref(self.instance.c_instance), sim.c_sim.n_forces, sim.c_sim.forces)
		self.n = self.c_system.n

	def __del__(self):
		try:
			self.binding.lib.bfm_system_destroy(C.byref(self.c_system))
		except Exception:
			pass

	def b(self) -> np.ndarray:
		return np.ctypeslib.as_array(self.c_system.b.data, shape=(self.n,)).copy()

	def get(self, i: int, j: int) -> float:
		return self.binding.lib.bfm_matrix_get(C.byref(self.c_system.A), i, j)

	def dense(self, copy: bool = True) -> np.ndarray:
		"""whole matrix through bfm_matrix_get (small systems only); copy=False: a view of a FULL row-major
		matrix's own storage (valid until the system changes), for systems too large to hold twice"""

		A = self.c_system.A

		if A.kind == abi.MATRIX_KIND_FULL and A.major == abi.MATRIX_MAJOR_ROW:
			view = np.ctypeslib.as_array(A.full.data, shape=(self.n, self.n))
			return view.copy() if copy else view

		get = self.binding.lib.bfm_matrix_get
		ref = C.byref(A)
		return np.array([[get(ref, i, j) for j in range(self.n)] for i in range(self.n)])

	def renumber(self):
		assert not self.binding.lib.bfm_system_renumber(C.byref(self.c_system))

	def perm(self):
		p = self.c_system.perm
		assert p.has_perm
		return (np.ctypeslib.as_array(p.perm, shape=(self.n,)).copy(), np.ctypeslib.as_array(p.inv_perm, shape=(self.n,)).copy())

	def bandwidth(self) -> int:
		return self.binding.lib.bfm_matrix_bandwidth(C.byref(self.c_system.A))

	def solve(self) -> np.ndarray:
		lib = self.binding.lib
		assert not lib.bfm_matrix_solve(C.byref(self.c_system.A), C.byref(self.c_system.b))

		if self.c_system.perm.has_perm:
			assert not lib.bfm_perm_perm_vec(C.byref(self.c_system.perm), C.byref(self.c_system.b), True)

		return self.b()
