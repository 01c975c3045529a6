/*
 * libbfm.h - the complete C ABI of libbfm (B200 build) in one translation-unit-friendly header.
 *
 * The reference spreads these declarations over 15 headers (libbfm/src/bfm/{bfm,math,matrix,mesh,
 * condition,force,material,shape,rule,obj,instance,sim,perm,system,ez}.h).  Here they live in one
 * file, grouped by layer; include/bfm/<name>.h are one-line forwarders so that sources written
 * against the reference (#include <bfm/sim.h> ...) and pybfm's cffi build compile unchanged.
 *
 * Struct layouts, enum values and the 67 function signatures are ABI-identical to the reference
 * (x86-64 sizes: state 64, vec 24, matrix 40, perm 40, system 120, mesh 88, edge 32, domain 72,
 * condition 40, force 48, material 72, shape 40, rule 88, obj 32, instance 48, sim 48, ez 368).
 * Every function returns 0 on success and -1 on failure unless stated otherwise.
 */
#ifndef BFM_LIBBFM_H
#define BFM_LIBBFM_H

#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ============================================================================================
 * bfm
 * ============================================================================================ */

// libbfm (B200 build) - library state: error slot + pluggable allocator.
// ABI-compatible with the reference's bfm/bfm.h:6-34 (sizeof(bfm_state_t) == 64 on x86-64).
// pybfm's binding generator (pybfm/bfm/gen_libbfm.py:11-50) takes its cdef text from the reference's
// own header files and only COMPILES against the installed bfm headers, so the forwarders suffice.

// allocator hooks; every buffer the library hands back is obtained through these
typedef void* (*bfm_alloc_t)(size_t size);
typedef void* (*bfm_realloc_t)(void* ptr, size_t size);
typedef void (*bfm_free_t)(void* ptr);

// error slot.  The reference never fills it; this build fills it when the GPU path fails
// (no device, CUDA error, CG did not converge) so that bfm_err_print() says why a call returned -1.
typedef struct {
	bool has;
	char* msg;

	char const* file;
	char const* func;
	size_t line;
} bfm_err_t;

typedef struct {
	bfm_err_t err;

	bfm_alloc_t alloc;
	bfm_realloc_t realloc;
	bfm_free_t free;
} bfm_state_t;

// all functions: 0 on success, -1 on failure (no errno)

int bfm_state_create(bfm_state_t* state);   // zeroes the state, installs malloc/realloc/free
int bfm_state_destroy(bfm_state_t* state);

int bfm_set_alloc(bfm_state_t* state, bfm_alloc_t alloc);
int bfm_set_realloc(bfm_state_t* state, bfm_realloc_t realloc);
int bfm_set_free(bfm_state_t* state, bfm_free_t free);

int bfm_err_print(bfm_state_t* state);      // prints "[BFM file:line (func)] msg" when err.has

/* ============================================================================================
 * math
 * ============================================================================================ */

// FP64 vector + numeric helpers.  ABI: reference bfm/math.h:5-23 (sizeof(bfm_vec_t) == 24).

#define BFM_NAN (0. / 0.)
#define BFM_IS_NAN(x) ((x) != (x))
#define BFM_PIVOT_EPS 1e-20
#define BFM_MAX(a, b) ((a) > (b) ? (a) : (b))
#define BFM_MIN(a, b) ((a) < (b) ? (a) : (b))
#define BFM_ABS(a) ((a) < 0 ? -(a) : (a))

typedef struct {
	bfm_state_t* state;

	size_t n;
	double* data; // host memory, state->alloc'd, zero-initialised by bfm_vec_create
} bfm_vec_t;

int bfm_vec_create(bfm_vec_t* vec, bfm_state_t* state, size_t n);
int bfm_vec_copy(bfm_vec_t* vec, bfm_vec_t* src); // copies src->n doubles; no size check (as the reference)
int bfm_vec_destroy(bfm_vec_t* vec);

/* ============================================================================================
 * matrix
 * ============================================================================================ */

// Square FP64 matrices.  ABI: reference bfm/matrix.h:10-127 (sizeof(bfm_matrix_t) == 40).
//
// FULL and BAND keep the reference's host storage and semantics.  CSR is this build's addition:
// a node-blocked sparse matrix living in GPU memory (the hot path's format); its union arm is an
// opaque handle that fits the existing 16-byte union, so the struct layout is unchanged.

typedef enum {
	BFM_MATRIX_KIND_FULL,
	BFM_MATRIX_KIND_BAND,
	BFM_MATRIX_KIND_CSR, // B200 build only: device-resident sparse matrix
} bfm_matrix_kind_t;

typedef enum {
	BFM_MATRIX_MAJOR_ROW,
	BFM_MATRIX_MAJOR_COLUMN,
} bfm_matrix_major_t;

// dense m*m doubles
typedef struct {
	double* data;
} bfm_matrix_full_t;

// band of half-width k: element (i, j) of a row-major matrix lives at data[j + i * 2k] inside an
// m * (2k + 1) buffer (reference matrix.c:196-247)
typedef struct {
	size_t k;
	double* data;
} bfm_matrix_band_t;

// opaque handle to the device-side sparse matrix (see include/bfm_b200.h)
typedef struct {
	void* impl;
	void* reserved;
} bfm_matrix_csr_t;

typedef struct {
	bfm_state_t* state;

	bfm_matrix_kind_t kind;
	bfm_matrix_major_t major;

	size_t m; // rows == columns

	union {
		bfm_matrix_full_t full;
		bfm_matrix_band_t band;
		bfm_matrix_csr_t csr;
	};
} bfm_matrix_t;

// zero-filled dense / band matrices (host)
int bfm_matrix_full_create(bfm_matrix_t* matrix, bfm_state_t* state, bfm_matrix_major_t major, size_t m);
int bfm_matrix_band_create(bfm_matrix_t* matrix, bfm_state_t* state, bfm_matrix_major_t major, size_t m, size_t k);

// same-kind copies are memcpy's, otherwise element by element; sizes must agree
int bfm_matrix_copy(bfm_matrix_t* matrix, bfm_matrix_t* src);
int bfm_matrix_destroy(bfm_matrix_t* matrix);

// element access; get returns NaN when (i, j) is out of range.  On a CSR matrix, get reads the
// value through a lazily refreshed host mirror and set/add only succeed inside the sparsity pattern.
double bfm_matrix_get(bfm_matrix_t* matrix, size_t i, size_t j);
int bfm_matrix_set(bfm_matrix_t* matrix, size_t i, size_t j, double val);
int bfm_matrix_add(bfm_matrix_t* matrix, size_t i, size_t j, double val);

// max |i - j| over the numerically non-zero entries (FULL, CSR) or the stored k (BAND)
size_t bfm_matrix_bandwidth(bfm_matrix_t* matrix);

// FULL/BAND: in-place unpivoted LU, then forward/backward substitution on y.
// CSR: bfm_matrix_lu prepares the Jacobi-scaled operator, bfm_matrix_lu_solve runs FP64
// preconditioned CG on the GPU and overwrites y with the solution.
int bfm_matrix_lu(bfm_matrix_t* matrix);
int bfm_matrix_lu_solve(bfm_matrix_t* matrix, bfm_vec_t* y);
int bfm_matrix_solve(bfm_matrix_t* matrix, bfm_vec_t* y);

/* ============================================================================================
 * mesh
 * ============================================================================================ */

// Mesh container and the two file readers.  ABI: reference bfm/mesh.h:7-54
// (sizeof: bfm_edge_t 32, bfm_domain_t 72, bfm_mesh_t 88).

typedef enum {
	BFM_ELEM_KIND_SIMPLEX = 3,            // P1 triangle
	BFM_ELEM_KIND_QUAD = 4,               // Q4 quadrilateral
	BFM_ELEM_KIND_QUADRATIC_TRIANGLE = 6, // declared by the reference, not supported by any solver path
} bfm_elem_kind_t;

typedef enum {
	BFM_PLANAR_STRAINS,
} bfm_problem_type_t;

// undirected edge nodes[0] <-> nodes[1]; elems[] are the adjacent elements, elems[1] == -1 on the boundary
typedef struct {
	size_t nodes[2];
	ssize_t elems[2];
} bfm_edge_t;

// named group of EDGE indices (LEPL1110 "domains")
typedef struct {
	char name[50];
	size_t n_elements;
	size_t* elements;
} bfm_domain_t;

typedef struct {
	bfm_state_t* state;

	size_t dim;
	bfm_elem_kind_t kind;

	size_t n_elems;
	size_t n_nodes;
	size_t n_edges;

	double* coords;    // [n_nodes][dim]
	size_t* elems;     // [n_elems][kind]
	bfm_edge_t* edges; // [n_edges]

	size_t n_domains;
	bfm_domain_t* domains;
} bfm_mesh_t;

int bfm_mesh_create(bfm_mesh_t* mesh, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind);
int bfm_mesh_destroy(bfm_mesh_t* mesh);

// LEPL1110 text format: nodes, edges, triangles|quads, domains (all four sections are required)
int bfm_mesh_read_lepl1110(bfm_mesh_t* mesh, bfm_state_t* state, char const* name);

// Wavefront OBJ: "v x y z" (z kept only when full), 1-based triangular "f a b c"; edges are derived
int bfm_mesh_read_wavefront(bfm_mesh_t* mesh, bfm_state_t* state, char const* name, bool full);

/* ============================================================================================
 * condition
 * ============================================================================================ */

// Boundary conditions.  ABI: reference bfm/condition.h:5-30 (sizeof(bfm_condition_t) == 40).

typedef enum {
	BFM_CONDITION_KIND_DIRICHLET_X = 0,
	BFM_CONDITION_KIND_DIRICHLET_Y = 1,
	BFM_CONDITION_KIND_NEUMANN_X = 2,
	BFM_CONDITION_KIND_NEUMANN_Y = 3,
	BFM_CONDITION_KIND_NEUMANN_NORMAL = 4,
	BFM_CONDITION_KIND_NEUMANN_TANGENT = 5,
	BFM_CONDITION_KIND_DIRICHLET_NORMAL = 6,
	BFM_CONDITION_KIND_DIRICHLET_TANGENT = 7,
} bfm_condition_kind_t;

typedef struct {
	bfm_state_t* state;
	bfm_mesh_t* mesh;

	bfm_condition_kind_t kind;
	double value;

	bool* nodes; // [mesh->n_nodes] membership mask, written directly by callers
} bfm_condition_t;

int bfm_condition_create(bfm_condition_t* condition, bfm_state_t* state, bfm_mesh_t* mesh, bfm_condition_kind_t kind);
int bfm_condition_destroy(bfm_condition_t* condition);

/* ============================================================================================
 * force
 * ============================================================================================ */

// Body forces.  ABI: reference bfm/force.h:6-46 (sizeof(bfm_force_t) == 48).

typedef enum {
	BFM_FORCE_KIND_NONE,
	BFM_FORCE_KIND_LINEAR, // constant vector
	BFM_FORCE_KIND_FUNKY,  // host callback; the GPU path samples it once per mesh node
} bfm_force_kind_t;

typedef struct bfm_force_t bfm_force_t;

typedef int (*bfm_force_funky_func_t)(bfm_force_t* force, bfm_vec_t* pos, bfm_vec_t* force_ref, void* data);

typedef struct {
	bfm_vec_t force;
} bfm_force_linear_t;

typedef struct {
	bfm_force_funky_func_t func;
	void* data;
} bfm_force_funky_t;

struct bfm_force_t {
	bfm_state_t* state;
	bfm_force_kind_t kind;
	size_t dim;

	union {
		bfm_force_linear_t linear;
		bfm_force_funky_t funky;
	};
};

int bfm_force_create(bfm_force_t* force, bfm_state_t* state, size_t dim);
int bfm_force_destroy(bfm_force_t* force);

int bfm_force_set_none(bfm_force_t* force);
int bfm_force_set_linear(bfm_force_t* force, bfm_vec_t* vec); // deep-copies vec; vec->n must equal dim
int bfm_force_set_funky(bfm_force_t* force, bfm_force_funky_func_t func, void* data);

// writes the force at pos into force_ref (force_ref->n must equal dim)
int bfm_force_eval(bfm_force_t* force, bfm_vec_t* pos, bfm_vec_t* force_ref);

/* ============================================================================================
 * material
 * ============================================================================================ */

// Isotropic linear-elastic material.  ABI: reference bfm/material.h:5-26 (sizeof(bfm_material_t) == 72).

typedef struct {
	double r;
	double g;
	double b;
	double a;
} bfm_colour_t;

typedef struct {
	bfm_state_t* state;

	char* name; // duplicated at creation
	bfm_colour_t colour;

	double rho; // density
	double E;   // Young's modulus
	double nu;  // Poisson's ratio
} bfm_material_t;

int bfm_material_create(bfm_material_t* material, bfm_state_t* state, char* name, double rho, double E, double nu);
int bfm_material_destroy(bfm_material_t* material);
int bfm_material_set_colour(bfm_material_t* material, double r, double g, double b, double a);

/* ============================================================================================
 * shape
 * ============================================================================================ */

// Shape functions.  ABI: reference bfm/shape.h:5-24 (sizeof(bfm_shape_t) == 40).
// The GPU assembly never calls these on the device: the host tabulates phi and its derivatives at
// the rule's integration points through these pointers and ships the table to the kernel.

typedef struct bfm_shape_t bfm_shape_t;

// values of the kind shape functions at a reference point
typedef int (*bfm_shape_fn_t)(bfm_shape_t* shape, double* point, double* phi);
// their derivatives with respect to reference coordinate wrt (0 = xsi, 1 = eta)
typedef int (*bfm_shape_dfn_t)(bfm_shape_t* shape, size_t wrt, double* point, double* dphi);

struct bfm_shape_t {
	bfm_state_t* state;

	size_t dim;
	bfm_elem_kind_t kind;

	bfm_shape_fn_t phi;
	bfm_shape_dfn_t dphi;
};

int bfm_shape_create(bfm_shape_t* shape, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind);
int bfm_shape_destroy(bfm_shape_t* shape);

/* ============================================================================================
 * rule
 * ============================================================================================ */

// Integration rules.  ABI: reference bfm/rule.h:6-24 (sizeof(bfm_rule_t) == 88).

typedef struct {
	bfm_state_t* state;

	size_t dim;
	bfm_elem_kind_t kind;
	size_t n_points;

	double* weights; // [n_points]
	double** points; // [n_points] -> [dim] reference coordinates

	bfm_shape_t shape;
} bfm_rule_t;

// generic rule with zeroed weights/points for the caller to fill
int bfm_rule_create(bfm_rule_t* rule, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind, size_t n_points);
int bfm_rule_destroy(bfm_rule_t* rule);

// 2-D Gauss-Legendre: 3 points on triangles, 2x2 points on quads
int bfm_rule_create_gauss_legendre(bfm_rule_t* rule, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind);

/* ============================================================================================
 * obj
 * ============================================================================================ */

// Simulation object = mesh + material + integration rule (borrowed pointers).
// ABI: reference bfm/obj.h:8-17 (sizeof(bfm_obj_t) == 32).

typedef struct {
	bfm_state_t* state;

	bfm_mesh_t* mesh;
	bfm_material_t* material;
	bfm_rule_t* rule;
} bfm_obj_t;

int bfm_obj_create(bfm_obj_t* obj, bfm_state_t* state, bfm_mesh_t* mesh, bfm_material_t* material, bfm_rule_t* rule);
int bfm_obj_destroy(bfm_obj_t* obj);

/* ============================================================================================
 * instance
 * ============================================================================================ */

// Instance = object + its boundary conditions + the solver's output.
// ABI: reference bfm/instance.h:6-23 (sizeof(bfm_instance_t) == 48).

typedef struct {
	bfm_state_t* state;
	bfm_obj_t* obj;

	size_t n_effects;
	double* effects; // [n_nodes][dim] displacements, host memory, written by bfm_sim_run

	size_t n_conditions;
	bfm_condition_t** conditions; // borrowed; must outlive bfm_sim_run
} bfm_instance_t;

int bfm_instance_create(bfm_instance_t* instance, bfm_state_t* state, bfm_obj_t* obj);
int bfm_instance_destroy(bfm_instance_t* instance);

int bfm_instance_set_n_conditions(bfm_instance_t* instance, size_t n_conditions);
int bfm_instance_add_condition(bfm_instance_t* instance, bfm_condition_t* condition);

/* ============================================================================================
 * sim
 * ============================================================================================ */

// Simulation driver.  ABI: reference bfm/sim.h:6-33 (sizeof(bfm_sim_t) == 48).

typedef enum {
	BFM_SIM_KIND_NONE = 0,
	BFM_SIM_KIND_PLANAR_STRAIN = 1,
	BFM_SIM_KIND_PLANAR_STRESS = 2,
	BFM_SIM_KIND_AXISYMMETRIC_STRAIN = 3,
} bfm_sim_kind_t;

typedef struct {
	bfm_state_t* state;
	bfm_sim_kind_t kind;

	size_t n_instances;
	bfm_instance_t** instances; // borrowed

	size_t n_forces;
	bfm_force_t** forces; // borrowed
} bfm_sim_t;

int bfm_sim_create(bfm_sim_t* sim, bfm_state_t* state, bfm_sim_kind_t kind);
int bfm_sim_destroy(bfm_sim_t* sim);

int bfm_sim_set_n_instances(bfm_sim_t* sim, size_t n_instances);
int bfm_sim_add_instance(bfm_sim_t* sim, bfm_instance_t* instance);

int bfm_sim_set_n_forces(bfm_sim_t* sim, size_t n_forces);
int bfm_sim_add_force(bfm_sim_t* sim, bfm_force_t* force);

// THE hot-path entry (reference sim.c:137-155): for every instance, assemble the elasticity
// system on the GPU, apply the boundary conditions, solve it with FP64 PCG and store the
// displacements in instance->effects.  Returns -1 if no CUDA device is usable - there is no CPU path.
int bfm_sim_run(bfm_sim_t* sim);

/* ============================================================================================
 * perm
 * ============================================================================================ */

// Permutations and Reverse Cuthill-McKee.  ABI: reference bfm/perm.h:6-22 (sizeof(bfm_perm_t) == 40).

typedef struct {
	bfm_state_t* state;

	size_t m;
	bool has_perm;

	size_t* perm;     // new index of old DOF i
	size_t* inv_perm; // old DOF sitting at new index i
} bfm_perm_t;

int bfm_perm_create(bfm_perm_t* perm, bfm_state_t* state, size_t m);
int bfm_perm_destroy(bfm_perm_t* perm);

// A'[p[i]][p[j]] = A[i][j] and v'[p[i]] = v[i], with p = inv ? inv_perm : perm
int bfm_perm_perm_matrix(bfm_perm_t* perm, bfm_matrix_t* matrix, bool inv);
int bfm_perm_perm_vec(bfm_perm_t* perm, bfm_vec_t* vec, bool inv);

// RCM on the numeric non-zero pattern of mat (FULL or CSR); bit-identical to the reference's ordering
int bfm_perm_rcm(bfm_perm_t* perm, bfm_matrix_t* mat);

/* ============================================================================================
 * system
 * ============================================================================================ */

// Linear system A x = b of one instance.  ABI: reference bfm/system.h:9-28 (sizeof(bfm_system_t) == 120).

typedef struct {
	bfm_state_t* state;

	size_t n;

	bfm_perm_t perm;
	bfm_matrix_t A;
	bfm_vec_t b;
} bfm_system_t;

// empty dense system (reference semantics: FULL row-major matrix, zero vector, no permutation)
int bfm_system_create(bfm_system_t* system, bfm_state_t* state, size_t n);
int bfm_system_destroy(bfm_system_t* system);

// RCM-renumber A and b.  FULL matrices become BAND as in the reference; CSR matrices stay CSR and
// are permuted logically (get/bandwidth/solve all act on the renumbered system).
int bfm_system_renumber(bfm_system_t* system);

// GPU assembly + boundary conditions; A comes back as a CSR matrix, b as a host vector
int bfm_system_create_planar_strain(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces);
int bfm_system_create_planar_stress(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces);
int bfm_system_create_axisymmetric_strain(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces);

/* ============================================================================================
 * ez
 * ============================================================================================ */

// LEPL1110 convenience layer: problem-file parser + U/V writer.
// ABI: reference bfm/ez.h:11-29 (sizeof(bfm_ez_lepl1110_t) == 368; the embedded structs are
// addressed directly by pybfm, pybfm/bfm/ez.py:18-22).

typedef struct {
	bfm_state_t* state;
	bfm_mesh_t* mesh;

	size_t n_conditions;
	bfm_condition_t* conditions;

	bfm_force_t gravity;
	bfm_material_t material;
	bfm_rule_t rule;
	bfm_obj_t obj;
	bfm_instance_t instance;
	bfm_sim_t sim;
} bfm_ez_lepl1110_t;

// the caller must hand in zeroed storage (ffi.new / calloc), as pybfm does
int bfm_ez_lepl1110_create(bfm_ez_lepl1110_t* ez, bfm_state_t* state, bfm_mesh_t* mesh, char* name);
int bfm_ez_lepl1110_destroy(bfm_ez_lepl1110_t* ez);

// writes component shift (0 = U, 1 = V) of the effects, "%14.7e", three per line
int bfm_ez_lepl1110_write(bfm_ez_lepl1110_t* ez, size_t shift, char const* filename);

#ifdef __cplusplus
}
#endif

#endif /* BFM_LIBBFM_H */
