/*
 * gpu.h - the thin C ABI between libbfm's host code (C) and its sm_100a CUDA kernels (.cu).
 *
 * Plain pointers and sizes only.  Every pointer named d_* is device memory obtained from
 * bfmg_alloc; everything runs on the library's own stream.  All functions return 0 on success and
 * -1 on failure (bfmg_last_error() says why).  There is no CPU fallback behind any of them.
 */
#ifndef BFM_GPU_H
#define BFM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- context --------------------------------------------------------------------------------- */

int bfmg_available(void);              /* 1 when a CUDA device is usable (initialises lazily) */
char const* bfmg_last_error(void);
int bfmg_sm_count(void);
size_t bfmg_launch_count(void);        /* kernels launched by this library so far */

int bfmg_alloc(void** d_ptr, size_t bytes);
void bfmg_free(void* d_ptr);
int bfmg_upload(void* d_dst, void const* src, size_t bytes);
int bfmg_download(void* dst, void const* d_src, size_t bytes); /* synchronises */
int bfmg_upload_narrow(int32_t* d_dst, size_t const* src, size_t count, size_t limit, int* out_of_range); /* size_t -> int32 on the way; flags values >= limit */
int bfmg_copy(void* d_dst, void const* d_src, size_t bytes);
int bfmg_host_pin(void const* ptr, size_t bytes);  /* page-lock a long-lived caller buffer in place (idempotent; -1: not pinned, copies still work) */
void bfmg_host_unpin(void const* ptr);             /* before the buffer is freed; harmless for buffers never pinned */
int bfmg_zero(void* d_dst, size_t bytes);
int bfmg_sync(void);

/* CUDA-event stopwatch on the library stream: tick returns a handle, lap gives ms between two */
int bfmg_tick(void);
float bfmg_lap(int from, int to);      /* synchronises on `to` */
int bfmg_timer_start(int slot);        /* 8 caller-owned stopwatch slots (same stream, CUDA events) */
float bfmg_timer_stop(int slot);       /* ms since the matching start; synchronises */

/* ---- sparsity pattern on the device (mirror of bfmi_plan_t, see internal.h) ------------------- */

typedef struct {
	int32_t nb;
	int32_t n_slices;
	int64_t n_slots;
	int32_t kind;

	int32_t* slice_off;
	int32_t* row_len;
	int32_t* scol;
	int32_t* diag_pos;
	int32_t* ctr_ptr;
	uint32_t* ctr;
	int32_t* elems;

	/* block rows the solver works on: [0, nb) on one GPU, the owned range of a partition otherwise
	 * (rows outside it are ghosts: gathered from, never computed) */
	int32_t row_lo, row_hi;
} bfmg_pattern_t;

/* ---- symbolic phase on the device (symbolic.cu): once per mesh ------------------------------------------- */

/* the SELL-32 node-block pattern and the element-to-nonzero map of a mesh from its connectivity (d_elems:
 * n_elems * kind node numbers, all < n_nodes; stays the caller's).  Fills nb, n_slices, n_slots, kind, row_lo / row_hi
 * and every array of *pat except elems (allocated here: bfmg_free); *n_ctr = entries of pat->ctr.  The arrays are
 * those plan.c's host builder produces, entry for entry. */
int bfmg_plan_build(int32_t n_nodes, int64_t n_elems, int32_t kind, int32_t const* d_elems, bfmg_pattern_t* pat, int64_t* n_ctr);

/* the edge list of reference mesh.c:32-102 (half-edges ordered by smaller node descending, larger ascending, ties in
 * element order; opposite neighbours fused): *d_edges = n_edges records of four int64 - nodes[0], nodes[1],
 * elems[0], elems[1] (-1 on the boundary), the layout of bfm_edge_t - allocated here (bfmg_free); NULL when empty */
int bfmg_edges_build(int32_t n_nodes, int64_t n_elems, int32_t kind, int32_t const* d_elems, int64_t** d_edges, int64_t* n_edges);

/* d_dst[i] = d_src[d_index[i]] for n node blocks (two doubles each): displacements from the internal numbering of
 * renumber.c back into the caller's */
int bfmg_gather_blocks(double* d_dst, double const* d_src, int32_t const* d_index, size_t n);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink (dist.cu) ------------------------------------ */

#define BFMG_DIST_ID_BYTES 128
#define BFMG_DIST_MAX_RANKS 16

int bfmg_dist_unique_id(void* id);                       /* BFMG_DIST_ID_BYTES bytes, made on one rank */
int bfmg_dist_init(int rank, int world, void const* id); /* collective over all ranks */
int bfmg_dist_finalize(void);
int bfmg_dist_world(void);                               /* 1 until bfmg_dist_init */
int bfmg_dist_rank(void);
size_t bfmg_dist_collectives(void);                      /* NCCL operations enqueued so far */
char const* bfmg_dist_p2p_status(void);                  /* "" when the NVLink peer-memory exchanges are in use, else why not */

/* halo plan of one rank (partition.c); arrays are host memory except d_send_idx */
typedef struct {
	int32_t n_nbr;
	int32_t const* nbr;        /* neighbour ranks, ascending */
	int32_t const* recv_begin; /* first ghost (local block row) owned by each neighbour */
	int32_t const* recv_count;
	int32_t const* send_ptr;   /* [n_nbr + 1] bounds in the packed send buffer */
	int32_t n_send;
	int32_t const* d_send_idx; /* device: local block rows to pack, grouped by neighbour */
} bfmg_halo_t;

/* refresh the ghost entries of d_vec (one double2 per local block row); d_sendbuf holds n_send double2 */
int bfmg_dist_halo(bfmg_halo_t const* halo, double* d_vec, double* d_sendbuf);
/* d_recv[world * count] = every rank's d_send[count], in rank order (a plain copy on one rank) */
int bfmg_dist_allgather_f64(double const* d_send, double* d_recv, int count);
/* d_global[2 * first_node[r] ...] = rank r's d_owned, for every r (first_node has world + 1 entries) */
int bfmg_dist_gather_blocks(double const* d_owned, size_t const* first_node, double* d_global);

/* ---- assembly -------------------------------------------------------------------------------- */

#define BFMG_MAX_POINTS 8
#define BFMG_MAX_FORCES 8

/* shape-function tables at the integration points, material constants, constant body forces */
typedef struct {
	int32_t kind;       /* nodes per element: 3 or 4 */
	int32_t n_points;
	int32_t axisym;     /* 0 planar, 1 axisymmetric */
	int32_t grad_const; /* 1 when dxsi/deta do not depend on the point (P1): geometry computed once */

	double a, b, c;     /* elasticity constants (reference system.c:454-458 / :242-244) */
	double rho;

	double weight[BFMG_MAX_POINTS];
	double phi[BFMG_MAX_POINTS][4];
	double dxsi[BFMG_MAX_POINTS][4];
	double deta[BFMG_MAX_POINTS][4];

	int32_t n_forces;
	int32_t forces_per_node; /* 0: const_force[k] applies everywhere; 1: d_nforce[k][node] tables */
	double const_force[BFMG_MAX_FORCES][2];
} bfmg_asm_tables_t;

/* one launch: stiffness blocks + load vector.  d_val: 4 * n_slots doubles, d_b: 2 * nb doubles.
 * Batch of independent systems in one pattern: d_tabs[system] (device) and d_slice_tab[slice] = system;
 * tab then only supplies the common kind / axisym.  NULL, NULL for a single system. */
int bfmg_assemble(bfmg_pattern_t const* pat, bfmg_asm_tables_t const* tab, double const* d_coords, double const* d_nforce, double* d_val, double* d_b, bfmg_asm_tables_t const* d_tabs, int32_t const* d_slice_tab);

/* ---- boundary conditions --------------------------------------------------------------------- */

/* one Dirichlet-type condition (reference apply_constraint, system.c:358-374, for every listed DOF
 * in ascending order): d_dofs/d_vals list the constrained DOFs, d_rows the block rows whose entries
 * or right-hand side can change; d_stamp/d_cval are 2*nb scratch arrays owned by the caller, epoch
 * must be unique per call and non-zero */
int bfmg_bc_dirichlet(bfmg_pattern_t const* pat, double* d_val, double* d_b, int32_t* d_stamp, double* d_cval, int32_t epoch, int32_t const* d_dofs, double const* d_vals, int32_t n_dofs, int32_t const* d_rows, int32_t n_rows);

/* right-hand-side additions (Neumann edge loads), grouped by DOF, applied in list order */
int bfmg_bc_add(double* d_b, int32_t const* d_group_dof, int32_t const* d_group_ptr, double const* d_add, int32_t n_groups);

/* ---- coarse level of the solver: rigid-body modes of node aggregates (coarse.c, coarse.cuh) --------- */

typedef struct {
	int32_t n_agg;
	int32_t n_colors;
	int32_t nc;          /* coarse dimension: 3 * n_agg rounded up to a multiple of 32 */
	int32_t half_bw;     /* E[i][j] = 0 for |i - j| > half_bw: adjacent aggregates have close numbers (bins in grid order) */

	int32_t* agg;        /* [nb] aggregate of every local block row (owned and ghost) */
	double* wgeom;       /* [nb] double2: node position relative to its aggregate's centroid */
	int32_t* agg_ptr;    /* [n_agg + 1] */
	int32_t* agg_nodes;  /* owned block rows of each aggregate, ascending */
	int32_t* color;      /* [n_agg] */
	int32_t* color_nbr;  /* [n_agg * n_colors] the aggregate of colour c equal or adjacent to g, or -1 */
} bfmg_coarse_t;

/* ---- multilevel preconditioner: aggregation hierarchy on the device (hier.c builds it, mg.cuh uses it) ----- */

#define BFMG_MG_MAX_LEVELS 12

typedef struct {
	int32_t n;           /* nodes of this level */
	int32_t dofs;        /* unknowns per node: 2 on level 0, 3 above */
	int32_t n_slices;
	int64_t n_slots;
	int32_t* slice_off;  /* SELL-32 node pattern of the level's operator */
	int32_t* scol;
	int32_t* diag_pos;
	int32_t* row_len;

	/* several GPUs (hier.c): rows [row_lo, row_hi) are this rank's; a distributed level refreshes its ghosts before
	 * every product with its operator; one GPU: row_lo = 0, row_hi = n */
	int32_t row_lo, row_hi;
	int32_t distributed;
	int32_t gather_first, gather_count; /* level below a replicated one: where this rank's aggregates sit in it */
	int32_t n_nbr, n_send;
	int32_t nbr[BFMG_DIST_MAX_RANKS];
	int32_t recv_begin[BFMG_DIST_MAX_RANKS];
	int32_t recv_count[BFMG_DIST_MAX_RANKS];
	int32_t send_ptr[BFMG_DIST_MAX_RANKS + 1];
	int32_t* send_idx;   /* device */

	/* towards the next level (unused on the last) */
	int32_t n_coarse;
	int32_t n_p;
	int32_t smoothed;    /* smoothed aggregation: rows that are all this rank's get (I - w A^) P~ (k_mg_smooth) */
	int32_t pad_;
	int32_t* agg;        /* [n] aggregate, -1: none */
	float* geom;         /* [n] float2 */
	int32_t* p_ptr;      /* prolongator entries by fine node */
	int32_t* p_col;
	int32_t* r_ptr;      /* ... and by coarse node */
	int32_t* r_ent;
	int32_t* r_node;
} bfmg_mg_level_t;

typedef struct {
	int32_t n_levels;    /* the last level is dense */
	int32_t nc;          /* its dimension: 3 * nodes rounded up to a multiple of 32 */
	int32_t half_bw;     /* band of the dense operator before inversion */
	bfmg_mg_level_t level[BFMG_MG_MAX_LEVELS];
} bfmg_mg_t;

/* ---- FP64 conjugate gradient ----------------------------------------------------------------- */

typedef struct {
	double tol;          /* stop at ||r|| <= tol * ||b|| in the Jacobi-scaled norm */
	int32_t max_iter;
	int32_t chunk;       /* iterations enqueued between two convergence polls */
	int32_t verify;      /* recompute the true residual b - A x at the end and report it */
	double true_tol;     /* a normwise backward error above this fails the solve (after max_restarts residual replacements) */
	int32_t max_restarts;
} bfmg_pcg_opts_t;

typedef struct {
	int32_t iterations;
	int32_t converged;   /* 1 converged, 0 hit max_iter, -1 breakdown */
	int32_t restarts;
	double rel_residual;      /* recursive residual, scaled norm */
	double true_rel_residual; /* ||b - A x|| / ||b||, scaled norm (NaN if not verified) */
	double backward_error;    /* ||b - A x|| / (||x|| + ||b||), scaled norm, ||A^|| taken as 1 (NaN if not verified) */
	float ms;
	float ms_setup;      /* of which: scaling + coarse-level setup (probing E, inverting it) */
	int32_t coarse_dim;  /* 0 when the coarse level was not used */
	int32_t mg_levels;   /* levels of the multilevel preconditioner (0: not used) */
	int32_t peer_memory; /* 1 when the exchanges of the iteration went over NVLink peer memory, 0: NCCL or one GPU */
	size_t launches;
} bfmg_pcg_result_t;

/* solves A x = b over the rows [pat->row_lo, pat->row_hi); d_val is left untouched (a scaled copy is
 * made).  halo = NULL on one GPU; otherwise the vectors are local (owned + ghost rows), every rank
 * calls this collectively and d_x receives the owned rows (ghost rows of d_x are not meaningful) */
int bfmg_pcg(bfmg_pattern_t const* pat, double const* d_val, double const* d_b, double* d_x, bfmg_pcg_opts_t const* opts, bfmg_pcg_result_t* res, bfmg_halo_t const* halo, bfmg_coarse_t const* coarse, bfmg_mg_t const* mg);

/* ---- small systems: one CTA per system (batch.cu) ------------------------------------------------ */

typedef struct {
	int32_t row_lo, row_hi; /* block rows of one system inside the shared pattern; row_lo % 32 == 0 */
} bfmg_batch_range_t;

typedef struct {
	int32_t iterations;
	int32_t converged;      /* 1 converged, 0 hit max_iter, -1 breakdown */
	double rel_residual;
	double true_rel_residual;
	double backward_error;
} bfmg_batch_status_t;

int bfmg_batch_max_rows(void); /* largest system (node rows) the one-CTA solver takes */

/* solves every system of the batch (ranges/status: host arrays of n_sys entries); systems that do not
 * converge are reported in status, not as a failure of the call */
int bfmg_pcg_batch(bfmg_pattern_t const* pat, double const* d_val, double const* d_b, double* d_x, bfmg_pcg_opts_t const* opts, int32_t n_sys, bfmg_batch_range_t const* ranges, bfmg_batch_status_t* status, float* ms);

/* times `reps` back-to-back launches of the CG SpMV kernel (q = A p with the fused dot) on d_val */
int bfmg_spmv_time(bfmg_pattern_t const* pat, double const* d_val, int reps, float* ms_per_launch);

/* y = A x, plain (tests) */
int bfmg_spmv(bfmg_pattern_t const* pat, double const* d_val, double const* d_x, double* d_y);

#ifdef __cplusplus
}
#endif

#endif
