"""ctypes driver of oracle/bfm_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbfm_oracle.so")

c_size_t_p = C.POINTER(C.c_size_t)
c_double_p = C.POINTER(C.c_double)


class OrcMesh(C.Structure):
	_fields_ = [
		("kind", C.c_int),
		("n_nodes", C.c_size_t),
		("n_elems", C.c_size_t),
		("n_edges", C.c_size_t),
		("coords", c_double_p),
		("elems", c_size_t_p),
		("edge_nodes", c_size_t_p),
		("edge_elems", C.POINTER(C.c_ssize_t)),
	]


class OrcCondition(C.Structure):
	_fields_ = [("kind", C.c_int), ("value", C.c_double), ("nodes", C.POINTER(C.c_uint8))]


class OrcProblem(C.Structure):
	_fields_ = [
		("mesh", OrcMesh),
		("sim_kind", C.c_int),
		("E", C.c_double),
		("nu", C.c_double),
		("rho", C.c_double),
		("n_points", C.c_size_t),
		("weights", c_double_p),
		("points", c_double_p),
		("n_forces", C.c_size_t),
		("forces", c_double_p),
		("n_conditions", C.c_size_t),
		("conditions", C.POINTER(OrcCondition)),
	]


class OrcSystem(C.Structure):
	_fields_ = [("n", C.c_size_t), ("rowptr", c_size_t_p), ("col", c_size_t_p), ("val", c_double_p), ("b", c_double_p)]


_lib = None


def build():
	subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)


def lib():
	global _lib

	if _lib is None:
		if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "bfm_oracle.c")):
			build()

		_lib = C.CDLL(LIB_PATH)
		P = C.POINTER
		_lib.orc_rule_gauss_legendre.argtypes = [C.c_int, c_double_p, c_double_p]
		_lib.orc_system_create.argtypes = [P(OrcProblem), P(OrcSystem)]
		_lib.orc_system_assemble.argtypes = [P(OrcProblem), P(OrcSystem)]
		_lib.orc_system_destroy.argtypes = [P(OrcSystem)]
		_lib.orc_system_destroy.restype = None
		_lib.orc_rcm.argtypes = [P(OrcSystem), c_size_t_p, c_size_t_p]
		_lib.orc_bandwidth.argtypes = [P(OrcSystem), c_size_t_p]
		_lib.orc_bandwidth.restype = C.c_size_t
		_lib.orc_band_solve.argtypes = [P(OrcSystem), c_size_t_p, c_double_p]
		_lib.orc_spmv.argtypes = [P(OrcSystem), c_double_p, c_double_p]
		_lib.orc_spmv.restype = None
		_lib.orc_run.argtypes = [P(OrcProblem), c_double_p]

	return _lib


def gauss_legendre(kind: int):
	w = np.zeros(kind)
	p = np.zeros((kind, 2))
	assert lib().orc_rule_gauss_legendre(kind, w.ctypes.data_as(c_double_p), p.ctypes.data_as(c_double_p)) == 0
	return w, p


class Problem:
	"""Flat description of one instance: mesh arrays + material + rule + forces + conditions.

	``conditions`` is a list of (kind, value, node mask); ``forces`` a list of either a constant
	(fx, fy) or an [n_nodes, 2] table (a FUNKY force sampled at the nodes)."""

	def __init__(self, coords, elems, sim_kind, E, nu, rho, forces=(), conditions=(), edges=None, rule=None):
		self.coords = np.ascontiguousarray(coords, dtype=np.float64)
		self.elems = np.ascontiguousarray(elems, dtype=np.uint64)
		self.kind = self.elems.shape[1]
		self.n_nodes = self.coords.shape[0]

		edges = np.zeros((0, 4), np.int64) if edges is None else np.asarray(edges, dtype=np.int64)
		self.edge_nodes = np.ascontiguousarray(edges[:, :2], dtype=np.uint64)
		self.edge_elems = np.ascontiguousarray(edges[:, 2:], dtype=np.int64)

		self.weights, self.points = rule if rule is not None else gauss_legendre(self.kind)
		self.weights = np.ascontiguousarray(self.weights, dtype=np.float64)
		self.points = np.ascontiguousarray(self.points, dtype=np.float64)

		tables = []

		for f in forces:
			f = np.asarray(f, dtype=np.float64)
			tables.append(np.broadcast_to(f, (self.n_nodes, 2)) if f.ndim == 1 else f)

		self.forces = np.ascontiguousarray(np.stack(tables) if tables else np.zeros((0, self.n_nodes, 2)))

		self.masks = [np.ascontiguousarray(m, dtype=np.uint8) for (_, _, m) in conditions]
		self.c_conditions = (OrcCondition * max(len(conditions), 1))()

		for i, (kind, value, _) in enumerate(conditions):
			self.c_conditions[i].kind = kind
			self.c_conditions[i].value = value
			self.c_conditions[i].nodes = self.masks[i].ctypes.data_as(C.POINTER(C.c_uint8))

		p = OrcProblem()
		p.mesh.kind = self.kind
		p.mesh.n_nodes = self.n_nodes
		p.mesh.n_elems = self.elems.shape[0]
		p.mesh.n_edges = self.edge_nodes.shape[0]
		p.mesh.coords = self.coords.ctypes.data_as(c_double_p)
		p.mesh.elems = self.elems.ctypes.data_as(c_size_t_p)
		p.mesh.edge_nodes = self.edge_nodes.ctypes.data_as(c_size_t_p)
		p.mesh.edge_elems = self.edge_elems.ctypes.data_as(C.POINTER(C.c_ssize_t))
		p.sim_kind = sim_kind
		p.E, p.nu, p.rho = E, nu, rho
		p.n_points = self.weights.shape[0]
		p.weights = self.weights.ctypes.data_as(c_double_p)
		p.points = self.points.ctypes.data_as(c_double_p)
		p.n_forces = self.forces.shape[0]
		p.forces = self.forces.ctypes.data_as(c_double_p)
		p.n_conditions = len(conditions)
		p.conditions = self.c_conditions
		self.c = p

	def run(self) -> np.ndarray:
		"""sim.c:103-135: displacements [n_nodes, 2]"""

		out = np.zeros(2 * self.n_nodes)
		assert lib().orc_run(C.byref(self.c), out.ctypes.data_as(c_double_p)) == 0
		return out.reshape(-1, 2)

	def system(self, with_bcs: bool = True) -> "System":
		return System(self, with_bcs)


class System:
	"""assembled sparse system (structural pattern, sorted columns)"""

	def __init__(self, problem: Problem, with_bcs: bool = True):
		self.problem = problem
		self.c = OrcSystem()
		fn = lib().orc_system_create if with_bcs else lib().orc_system_assemble
		assert fn(C.byref(problem.c), C.byref(self.c)) == 0

		self.n = self.c.n
		self.rowptr = np.ctypeslib.as_array(self.c.rowptr, shape=(self.n + 1,))
		nnz = int(self.rowptr[-1])
		self.col = np.ctypeslib.as_array(self.c.col, shape=(nnz,))
		self.val = np.ctypeslib.as_array(self.c.val, shape=(nnz,))
		self.b = np.ctypeslib.as_array(self.c.b, shape=(self.n,))

	def __del__(self):
		try:
			lib().orc_system_destroy(C.byref(self.c))
		except Exception:
			pass

	def dense(self) -> np.ndarray:
		A = np.zeros((self.n, self.n))
		rows = np.repeat(np.arange(self.n), np.diff(self.rowptr).astype(np.int64))
		A[rows, self.col.astype(np.int64)] = self.val
		return A

	def scipy(self):
		import scipy.sparse as sp

		return sp.csr_matrix((self.val.copy(), self.col.astype(np.int64), self.rowptr.astype(np.int64)), shape=(self.n, self.n))

	def rcm(self):
		perm = np.zeros(self.n, np.uint64)
		inv = np.zeros(self.n, np.uint64)
		assert lib().orc_rcm(C.byref(self.c), perm.ctypes.data_as(c_size_t_p), inv.ctypes.data_as(c_size_t_p)) == 0
		return perm, inv

	def bandwidth(self, perm=None) -> int:
		arg = perm.ctypes.data_as(c_size_t_p) if perm is not None else None
		return lib().orc_bandwidth(C.byref(self.c), arg)

	def band_solve(self, perm=None) -> np.ndarray:
		x = np.zeros(self.n)
		arg = perm.ctypes.data_as(c_size_t_p) if perm is not None else None
		assert lib().orc_band_solve(C.byref(self.c), arg, x.ctypes.data_as(c_double_p)) == 0
		return x

	def spmv(self, x) -> np.ndarray:
		x = np.ascontiguousarray(x, dtype=np.float64)
		y = np.zeros(self.n)
		lib().orc_spmv(C.byref(self.c), x.ctypes.data_as(c_double_p), y.ctypes.data_as(c_double_p))
		return y
