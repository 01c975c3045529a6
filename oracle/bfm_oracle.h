/*
 * bfm_oracle.h - CPU restatement of libbfm's FEM hot path.
 *
 * TEST INFRASTRUCTURE.  Nothing in the shipped library (bfm_b200/) includes, links or calls this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it,
 * and only as the checker or the timed CPU baseline.
 *
 * The reference (obiwac/bfm, libbfm) stores the global matrix as a dense n*n array, which caps it
 * at n <= 46 340.  This restatement keeps the reference's arithmetic - the same loops in the same
 * order, the same expression association, no fused multiply-add - but keeps the matrix in a sorted
 * sparse row structure, so that it stays bit-identical to the reference where the reference can run
 * and remains usable (assembly, BCs, RCM, band LU) well beyond it.
 *
 * Parity pin: tests/test_oracle.py checks this port bit-for-bit against the reference compiled
 * from its own sources (oracle/_ref/libbfm_ref.so, see oracle/Makefile) and against the reference's
 * golden vectors data/U.txt, data/V.txt (committed as tests/golden/lepl8_{U,V}.txt).
 */
#ifndef BFM_ORACLE_H
#define BFM_ORACLE_H

#include <stddef.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mesh: what bfm_mesh_t carries (bfm/mesh.h:31-48), flattened */
typedef struct {
	int kind;                 /* 3 = P1 triangle, 4 = Q4 quad (bfm/mesh.h:7-11) */
	size_t n_nodes;
	size_t n_elems;
	size_t n_edges;
	const double* coords;     /* [n_nodes][2] */
	const size_t* elems;      /* [n_elems][kind] */
	const size_t* edge_nodes; /* [n_edges][2] */
	const ssize_t* edge_elems;/* [n_edges][2], [1] == -1 on the boundary */
} orc_mesh_t;

/* one boundary condition (bfm/condition.h:5-27) */
typedef struct {
	int kind;                   /* bfm_condition_kind_t value, 0..7 */
	double value;
	const unsigned char* nodes; /* [n_nodes] mask */
} orc_condition_t;

typedef struct {
	orc_mesh_t mesh;

	int sim_kind;             /* 1 planar strain, 2 planar stress, 3 axisymmetric (bfm/sim.h:6-11) */
	double E, nu, rho;        /* bfm/material.h:12-21 */

	size_t n_points;          /* integration rule (bfm/rule.h:6-17) */
	const double* weights;    /* [n_points] */
	const double* points;     /* [n_points][2] */

	/* body forces sampled at the nodes: the reference only ever evaluates a force at element node
	 * positions (system.c:190-203), so a per-node table covers NONE, LINEAR and FUNKY alike */
	size_t n_forces;
	const double* forces;     /* [n_forces][n_nodes][2] */

	size_t n_conditions;
	const orc_condition_t* conditions;
} orc_problem_t;

/* assembled system: rows hold the STRUCTURAL pattern (all DOF pairs sharing an element), sorted by
 * column; entries may be numerically zero exactly as in the reference's dense matrix */
typedef struct {
	size_t n;
	size_t* rowptr;           /* [n+1] */
	size_t* col;              /* [nnz] */
	double* val;              /* [nnz] */
	double* b;                /* [n] */
} orc_system_t;

/* rule.c:90-111 */
int orc_rule_gauss_legendre(int kind, double* weights, double* points);

/* system.c:427-529 (planar) and :539-622 (axisymmetric): assembly followed by the BC loop */
int orc_system_create(const orc_problem_t* problem, orc_system_t* out);
/* same, stopping before the BC loop (for per-stage parity tests) */
int orc_system_assemble(const orc_problem_t* problem, orc_system_t* out);
void orc_system_destroy(orc_system_t* sys);

/* dense read-back A[i][j] (0 outside the structural pattern) */
double orc_system_get(const orc_system_t* sys, size_t i, size_t j);

/* perm.c:117-331 on the numeric nonzero pattern */
int orc_rcm(const orc_system_t* sys, size_t* perm, size_t* inv_perm);

/* matrix.c:57-71 evaluated on the permuted matrix (perm may be NULL for identity) */
size_t orc_bandwidth(const orc_system_t* sys, const size_t* perm);

/* system.c:44-81 (permute, to band) + matrix.c:253-302, 351-404 (band LU, substitution) +
 * sim.c:123 (inverse permutation).  x receives the solution in the ORIGINAL numbering.
 * Returns 0, or -1 where the reference's bfm_matrix_solve would fail. */
int orc_band_solve(const orc_system_t* sys, const size_t* perm, double* x);

/* y = A x with the sparse structure (used by tests for true residuals) */
void orc_spmv(const orc_system_t* sys, const double* x, double* y);

/* sim.c:103-135 for one instance: create, renumber, solve, un-permute -> effects[n] */
int orc_run(const orc_problem_t* problem, double* effects);

#ifdef __cplusplus
}
#endif

#endif
