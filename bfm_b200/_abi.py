"""ctypes view of libbfm's C ABI (struct layouts + the 67 exported functions).

The layouts restate the reference headers ``libbfm/src/bfm/*.h`` (x86-64 sizes checked against
SURVEY.md section 8b: state 64, vec 24, matrix 40, perm 40, system 120, mesh 88, edge 32, domain 72,
condition 40, force 48, material 72, shape 40, rule 88, obj 32, instance 48, sim 48, ez 368).
Because the ABI is shared, the same declarations bind either our ``libbfm.so`` or the reference
library compiled by ``oracle/Makefile`` - the parity tests drive both through this one module.
"""

import ctypes as C
import os

c_size_t = C.c_size_t
c_ssize_t = C.c_ssize_t
c_double_p = C.POINTER(C.c_double)
c_size_t_p = C.POINTER(c_size_t)

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, c_size_t)
REALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, c_size_t)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p)


class Err(C.Structure):  # bfm/bfm.h:6-13
	_fields_ = [("has", C.c_bool), ("msg", C.c_char_p), ("file", C.c_char_p), ("func", C.c_char_p), ("line", c_size_t)]


class State(C.Structure):  # bfm/bfm.h:19-25
	_fields_ = [("err", Err), ("alloc", ALLOC_FN), ("realloc", REALLOC_FN), ("free", FREE_FN)]


class Vec(C.Structure):  # bfm/math.h:13-18
	_fields_ = [("state", C.POINTER(State)), ("n", c_size_t), ("data", c_double_p)]


class MatrixFull(C.Structure):  # bfm/matrix.h:20-22
	_fields_ = [("data", c_double_p)]


class MatrixBand(C.Structure):  # bfm/matrix.h:24-27
	_fields_ = [("k", c_size_t), ("data", c_double_p)]


class MatrixCsr(C.Structure):  # our third arm: fits the 16-byte union (include/bfm/matrix.h)
	_fields_ = [("impl", C.c_void_p), ("reserved", C.c_void_p)]


class _MatrixU(C.Union):
	_fields_ = [("full", MatrixFull), ("band", MatrixBand), ("csr", MatrixCsr)]


class Matrix(C.Structure):  # bfm/matrix.h:29-42
	_anonymous_ = ("u",)
	_fields_ = [("state", C.POINTER(State)), ("kind", C.c_int), ("major", C.c_int), ("m", c_size_t), ("u", _MatrixU)]


MATRIX_KIND_FULL, MATRIX_KIND_BAND, MATRIX_KIND_CSR = 0, 1, 2
MATRIX_MAJOR_ROW, MATRIX_MAJOR_COLUMN = 0, 1


class Edge(C.Structure):  # bfm/mesh.h:17-23
	_fields_ = [("nodes", c_size_t * 2), ("elems", c_ssize_t * 2)]


class Domain(C.Structure):  # bfm/mesh.h:25-29
	_fields_ = [("name", C.c_char * 50), ("n_elements", c_size_t), ("elements", c_size_t_p)]


class Mesh(C.Structure):  # bfm/mesh.h:31-48
	_fields_ = [
		("state", C.POINTER(State)),
		("dim", c_size_t),
		("kind", C.c_int),
		("n_elems", c_size_t),
		("n_nodes", c_size_t),
		("n_edges", c_size_t),
		("coords", c_double_p),
		("elems", c_size_t_p),
		("edges", C.POINTER(Edge)),
		("n_domains", c_size_t),
		("domains", C.POINTER(Domain)),
	]


class Condition(C.Structure):  # bfm/condition.h:16-27
	_fields_ = [
		("state", C.POINTER(State)),
		("mesh", C.POINTER(Mesh)),
		("kind", C.c_int),
		("value", C.c_double),
		("nodes", C.POINTER(C.c_bool)),
	]


class ForceLinear(C.Structure):  # bfm/force.h:12-14
	_fields_ = [("force", Vec)]


class ForceFunky(C.Structure):  # bfm/force.h:20-23
	_fields_ = [("func", C.c_void_p), ("data", C.c_void_p)]


class _ForceU(C.Union):
	_fields_ = [("linear", ForceLinear), ("funky", ForceFunky)]


class Force(C.Structure):  # bfm/force.h:25-34
	_anonymous_ = ("u",)
	_fields_ = [("state", C.POINTER(State)), ("kind", C.c_int), ("dim", c_size_t), ("u", _ForceU)]


FORCE_FUNKY_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Force), C.POINTER(Vec), C.POINTER(Vec), C.c_void_p)  # bfm/force.h:18


class Colour(C.Structure):  # bfm/material.h:5-10
	_fields_ = [("r", C.c_double), ("g", C.c_double), ("b", C.c_double), ("a", C.c_double)]


class Material(C.Structure):  # bfm/material.h:12-21
	_fields_ = [
		("state", C.POINTER(State)),
		("name", C.c_void_p),
		("colour", Colour),
		("rho", C.c_double),
		("E", C.c_double),
		("nu", C.c_double),
	]


class Shape(C.Structure):  # bfm/shape.h:13-21
	_fields_ = [("state", C.POINTER(State)), ("dim", c_size_t), ("kind", C.c_int), ("phi", C.c_void_p), ("dphi", C.c_void_p)]


class Rule(C.Structure):  # bfm/rule.h:6-17
	_fields_ = [
		("state", C.POINTER(State)),
		("dim", c_size_t),
		("kind", C.c_int),
		("n_points", c_size_t),
		("weights", c_double_p),
		("points", C.POINTER(c_double_p)),
		("shape", Shape),
	]


class Obj(C.Structure):  # bfm/obj.h:8-14
	_fields_ = [("state", C.POINTER(State)), ("mesh", C.POINTER(Mesh)), ("material", C.POINTER(Material)), ("rule", C.POINTER(Rule))]


class Instance(C.Structure):  # bfm/instance.h:6-17
	_fields_ = [
		("state", C.POINTER(State)),
		("obj", C.POINTER(Obj)),
		("n_effects", c_size_t),
		("effects", c_double_p),
		("n_conditions", c_size_t),
		("conditions", C.POINTER(C.POINTER(Condition))),
	]


class Sim(C.Structure):  # bfm/sim.h:13-22
	_fields_ = [
		("state", C.POINTER(State)),
		("kind", C.c_int),
		("n_instances", c_size_t),
		("instances", C.POINTER(C.POINTER(Instance))),
		("n_forces", c_size_t),
		("forces", C.POINTER(C.POINTER(Force))),
	]


class Perm(C.Structure):  # bfm/perm.h:6-14
	_fields_ = [("state", C.POINTER(State)), ("m", c_size_t), ("has_perm", C.c_bool), ("perm", c_size_t_p), ("inv_perm", c_size_t_p)]


class System(C.Structure):  # bfm/system.h:9-17
	_fields_ = [("state", C.POINTER(State)), ("n", c_size_t), ("perm", Perm), ("A", Matrix), ("b", Vec)]


class Ez(C.Structure):  # bfm/ez.h:11-24
	_fields_ = [
		("state", C.POINTER(State)),
		("mesh", C.POINTER(Mesh)),
		("n_conditions", c_size_t),
		("conditions", C.POINTER(Condition)),
		("gravity", Force),
		("material", Material),
		("rule", Rule),
		("obj", Obj),
		("instance", Instance),
		("sim", Sim),
	]


EXPECTED_SIZES = {
	State: 64, Vec: 24, Matrix: 40, Perm: 40, System: 120, Mesh: 88, Edge: 32, Domain: 72, Condition: 40,
	Force: 48, Material: 72, Shape: 40, Rule: 88, Obj: 32, Instance: 48, Sim: 48, Ez: 368,
}

_P = C.POINTER
_int = C.c_int

# name -> (restype, argtypes); the 67 functions the reference headers declare
PROTOTYPES = {
	# bfm/bfm.h:27-34
	"bfm_state_create": (_int, [_P(State)]),
	"bfm_state_destroy": (_int, [_P(State)]),
	"bfm_set_alloc": (_int, [_P(State), ALLOC_FN]),
	"bfm_set_realloc": (_int, [_P(State), REALLOC_FN]),
	"bfm_set_free": (_int, [_P(State), FREE_FN]),
	"bfm_err_print": (_int, [_P(State)]),
	# bfm/math.h:21-23
	"bfm_vec_create": (_int, [_P(Vec), _P(State), c_size_t]),
	"bfm_vec_copy": (_int, [_P(Vec), _P(Vec)]),
	"bfm_vec_destroy": (_int, [_P(Vec)]),
	# bfm/matrix.h:54-127
	"bfm_matrix_full_create": (_int, [_P(Matrix), _P(State), _int, c_size_t]),
	"bfm_matrix_band_create": (_int, [_P(Matrix), _P(State), _int, c_size_t, c_size_t]),
	"bfm_matrix_copy": (_int, [_P(Matrix), _P(Matrix)]),
	"bfm_matrix_destroy": (_int, [_P(Matrix)]),
	"bfm_matrix_get": (C.c_double, [_P(Matrix), c_size_t, c_size_t]),
	"bfm_matrix_set": (_int, [_P(Matrix), c_size_t, c_size_t, C.c_double]),
	"bfm_matrix_add": (_int, [_P(Matrix), c_size_t, c_size_t, C.c_double]),
	"bfm_matrix_bandwidth": (c_size_t, [_P(Matrix)]),
	"bfm_matrix_lu": (_int, [_P(Matrix)]),
	"bfm_matrix_lu_solve": (_int, [_P(Matrix), _P(Vec)]),
	"bfm_matrix_solve": (_int, [_P(Matrix), _P(Vec)]),
	# bfm/mesh.h:50-54
	"bfm_mesh_create": (_int, [_P(Mesh), _P(State), c_size_t, _int]),
	"bfm_mesh_destroy": (_int, [_P(Mesh)]),
	"bfm_mesh_read_lepl1110": (_int, [_P(Mesh), _P(State), C.c_char_p]),
	"bfm_mesh_read_wavefront": (_int, [_P(Mesh), _P(State), C.c_char_p, C.c_bool]),
	# bfm/condition.h:29-30
	"bfm_condition_create": (_int, [_P(Condition), _P(State), _P(Mesh), _int]),
	"bfm_condition_destroy": (_int, [_P(Condition)]),
	# bfm/force.h:36-46
	"bfm_force_create": (_int, [_P(Force), _P(State), c_size_t]),
	"bfm_force_destroy": (_int, [_P(Force)]),
	"bfm_force_set_none": (_int, [_P(Force)]),
	"bfm_force_set_linear": (_int, [_P(Force), _P(Vec)]),
	"bfm_force_set_funky": (_int, [_P(Force), FORCE_FUNKY_FN, C.c_void_p]),
	"bfm_force_eval": (_int, [_P(Force), _P(Vec), _P(Vec)]),
	# bfm/material.h:23-26
	"bfm_material_create": (_int, [_P(Material), _P(State), C.c_char_p, C.c_double, C.c_double, C.c_double]),
	"bfm_material_destroy": (_int, [_P(Material)]),
	"bfm_material_set_colour": (_int, [_P(Material), C.c_double, C.c_double, C.c_double, C.c_double]),
	# bfm/rule.h:19-24
	"bfm_rule_create": (_int, [_P(Rule), _P(State), c_size_t, _int, c_size_t]),
	"bfm_rule_destroy": (_int, [_P(Rule)]),
	"bfm_rule_create_gauss_legendre": (_int, [_P(Rule), _P(State), c_size_t, _int]),
	# bfm/shape.h:23-24
	"bfm_shape_create": (_int, [_P(Shape), _P(State), c_size_t, _int]),
	"bfm_shape_destroy": (_int, [_P(Shape)]),
	# bfm/obj.h:16-17
	"bfm_obj_create": (_int, [_P(Obj), _P(State), _P(Mesh), _P(Material), _P(Rule)]),
	"bfm_obj_destroy": (_int, [_P(Obj)]),
	# bfm/instance.h:19-23
	"bfm_instance_create": (_int, [_P(Instance), _P(State), _P(Obj)]),
	"bfm_instance_destroy": (_int, [_P(Instance)]),
	"bfm_instance_set_n_conditions": (_int, [_P(Instance), c_size_t]),
	"bfm_instance_add_condition": (_int, [_P(Instance), _P(Condition)]),
	# bfm/sim.h:24-33
	"bfm_sim_create": (_int, [_P(Sim), _P(State), _int]),
	"bfm_sim_destroy": (_int, [_P(Sim)]),
	"bfm_sim_set_n_instances": (_int, [_P(Sim), c_size_t]),
	"bfm_sim_add_instance": (_int, [_P(Sim), _P(Instance)]),
	"bfm_sim_set_n_forces": (_int, [_P(Sim), c_size_t]),
	"bfm_sim_add_force": (_int, [_P(Sim), _P(Force)]),
	"bfm_sim_run": (_int, [_P(Sim)]),
	# bfm/perm.h:16-22
	"bfm_perm_create": (_int, [_P(Perm), _P(State), c_size_t]),
	"bfm_perm_destroy": (_int, [_P(Perm)]),
	"bfm_perm_perm_matrix": (_int, [_P(Perm), _P(Matrix), C.c_bool]),
	"bfm_perm_perm_vec": (_int, [_P(Perm), _P(Vec), C.c_bool]),
	"bfm_perm_rcm": (_int, [_P(Perm), _P(Matrix)]),
	# bfm/system.h:19-28
	"bfm_system_create": (_int, [_P(System), _P(State), c_size_t]),
	"bfm_system_destroy": (_int, [_P(System)]),
	"bfm_system_renumber": (_int, [_P(System)]),
	"bfm_system_create_planar_strain": (_int, [_P(System), _P(Instance), c_size_t, _P(_P(Force))]),
	"bfm_system_create_planar_stress": (_int, [_P(System), _P(Instance), c_size_t, _P(_P(Force))]),
	"bfm_system_create_axisymmetric_strain": (_int, [_P(System), _P(Instance), c_size_t, _P(_P(Force))]),
	# bfm/ez.h:26-29
	"bfm_ez_lepl1110_create": (_int, [_P(Ez), _P(State), _P(Mesh), C.c_char_p]),
	"bfm_ez_lepl1110_destroy": (_int, [_P(Ez)]),
	"bfm_ez_lepl1110_write": (_int, [_P(Ez), c_size_t, C.c_char_p]),
}

assert len(PROTOTYPES) == 67


def bind(path: str, extra: dict | None = None) -> C.CDLL:
	"""dlopen ``path`` and attach prototypes; raises if any of the 67 symbols is missing."""

	lib = C.CDLL(os.path.abspath(path), mode=C.RTLD_LOCAL)  # never global: ours and the reference export the same names

	for name, (restype, argtypes) in {**PROTOTYPES, **(extra or {})}.items():
		fn = getattr(lib, name)  # AttributeError = missing export
		fn.restype = restype
		fn.argtypes = argtypes

	return lib


for _cls, _size in EXPECTED_SIZES.items():
	assert C.sizeof(_cls) == _size, (_cls.__name__, C.sizeof(_cls), _size)
