/*
 * bfm_b200.h - additions of the B200 build of libbfm, on top of the reference-compatible C API in
 * <bfm/libbfm.h>.  All symbols are prefixed bfmx_.  C ABI: plain pointers and sizes.
 *
 * What is here and why:
 *   - bfmx_job_*      the stages bfm_sim_run (reference sim.c:103-135) runs for one instance, exposed
 *                     one by one so a caller (bench.py) can keep inputs resident in HBM and time
 *                     assembly and solve separately with the library's own CUDA events;
 *   - bfmx_stats_t    per-run counters and device timings (the reference reports nothing);
 *   - bfmx_mesh_*     in-memory mesh helpers: the reference can only build meshes by reading files;
 *   - bfmx_matrix_csr_create  hand a sparse system of your own to the GPU solver through the
 *                     ordinary bfm_matrix_* / bfm_perm_* calls.
 */
#ifndef BFM_B200_H
#define BFM_B200_H

#include <stdint.h>

#include <bfm/libbfm.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 1 when a CUDA device is usable.  Without one every hot-path call returns -1: there is no CPU path. */
int bfmx_device_available(void);
/* message of the last device-side failure in this process ("" if none) */
char const* bfmx_device_error(void);

/* wait for everything the library has queued on its stream */
int bfmx_device_sync(void);
/* CUDA-event stopwatch on the library's stream (8 slots): stop returns the milliseconds since start */
int bfmx_timer_start(int slot);
float bfmx_timer_stop(int slot);
/* kernels this library has launched so far in this process */
size_t bfmx_kernel_launches(void);
int bfmx_device_sm_count(void);

typedef struct {
	size_t n_dofs;
	size_t n_blocks;           /* 2x2 node blocks of the structural pattern */
	size_t n_slots;            /* blocks actually stored (SELL-32 padding included) */

	int cg_iterations;
	int cg_restarts;           /* residual replacements triggered by the true-residual check */
	int cg_converged;          /* 1 yes, 0 iteration limit / residual drift, -1 breakdown */
	double cg_rel_residual;      /* recursive ||r|| / ||b||, Jacobi-scaled norm */
	double cg_true_rel_residual; /* recomputed ||b - A x|| / ||b||, same norm */

	/* device times, CUDA events on the library stream (ms); plan is host wall-clock */
	float ms_plan;             /* symbolic phase + its upload (0 when the cached plan was reused) */
	float ms_upload;           /* coordinates, force tables, BC lists: host -> device */
	float ms_assemble;         /* assembly kernel */
	float ms_bc;               /* boundary-condition kernels */
	float ms_solve;            /* whole PCG, including setup kernels */
	float ms_download;         /* displacements: device -> host */

	size_t kernel_launches;
	size_t h2d_bytes;
	size_t d2h_bytes;

	/* recomputed ||b - A x|| / (||x|| + ||b||) in the Jacobi-scaled norm (||A^|| >= 1 taken as 1): the
	 * normwise backward error, which is what FP64 can guarantee on an ill-conditioned system */
	double cg_backward_error;

	/* multi-GPU runs (bfmx_dist_init): this rank's share; n_dofs stays the global size */
	size_t n_dofs_owned;
	size_t n_ranks;
	size_t halo_bytes_per_exchange;

	size_t coarse_dim;         /* dimension of the solver's coarse space (3 rigid-body modes per aggregate); 0 = none */
	float ms_solve_setup;      /* part of ms_solve spent scaling the matrix and building the coarse operator */
	int uses_peer_memory;      /* multi-GPU: 1 when the per-iteration exchanges went over NVLink peer memory (CUDA IPC), 0 when over NCCL */
	int mg_levels;             /* levels of the solver's aggregation multigrid preconditioner (mesh level and dense last level included); 0 = not used */
} bfmx_stats_t;

/* stats of the most recent bfm_sim_run instance / bfm_matrix_solve / job stage in this process */
int bfmx_last_stats(bfmx_stats_t* out);
void bfmx_publish_stats(bfmx_stats_t const* stats);

/* ---- staged pipeline for one instance of a simulation -------------------------------------------- */

typedef struct bfmx_job bfmx_job_t;

/* validates the instance, builds or reuses the symbolic plan of its mesh, tabulates the shape
 * functions at the rule's points, turns the boundary conditions into ordered device work lists */
int bfmx_job_create(bfmx_job_t** job, bfm_sim_t* sim, size_t instance_index);
int bfmx_job_upload(bfmx_job_t* job);    /* inputs host -> device */
int bfmx_job_assemble(bfmx_job_t* job);  /* assembly + boundary conditions, device only */
int bfmx_job_solve(bfmx_job_t* job);     /* FP64 PCG, device only (plus a 64-byte status poll per chunk) */
int bfmx_job_download(bfmx_job_t* job);  /* displacements -> instance->effects */
int bfmx_job_stats(bfmx_job_t* job, bfmx_stats_t* out);
/* average duration of `reps` back-to-back launches of the CG SpMV kernel on the assembled matrix */
int bfmx_job_spmv_time(bfmx_job_t* job, int reps, float* ms_per_launch);
/* copies of the assembled right-hand side / solution (n_dofs doubles each; tests) */
int bfmx_job_read(bfmx_job_t* job, double* b_or_null, double* x_or_null);
int bfmx_job_destroy(bfmx_job_t* job);

/* ---- batches of small systems (BASELINE.json configs[4]) ---------------------------------------------
 *
 * Every instance of every simulation becomes one independent system; all of them are assembled by one
 * launch and each is solved entirely by its own CTA (bfm_b200/csrc/batch.cu).  All systems must use the
 * same element kind and must not mix planar with axisymmetric problems; each needs at most
 * bfmx_batch_max_nodes() nodes.  A batch job goes through the same stages as a single one
 * (bfmx_job_upload / _assemble / _solve / _download / _stats / _destroy). */

typedef struct {
	int iterations;
	int converged;             /* 1 yes, 0 iteration limit or backward error above tolerance, -1 breakdown */
	double rel_residual;
	double true_rel_residual;
	double backward_error;
} bfmx_batch_status_t;

int bfmx_batch_max_nodes(void);
int bfmx_job_create_batch(bfmx_job_t** job, bfm_sim_t** sims, size_t n_sims);
int bfmx_job_batch_size(bfmx_job_t* job);
int bfmx_job_batch_status(bfmx_job_t* job, size_t system, bfmx_batch_status_t* out);
/* create + upload + assemble + solve + download + destroy: bfm_sim_run for n_sims simulations at once */
int bfmx_sim_run_batch(bfm_sim_t** sims, size_t n_sims);

/* ---- meshes ------------------------------------------------------------------------------------------ */

/* (re)derive mesh->edges from the connectivity exactly as the Wavefront reader does (reference mesh.c:52-102) */
int bfmx_mesh_compute_edges(bfm_mesh_t* mesh);

/* Does bfm_sim_run work on an internally renumbered copy of this mesh (bfm_b200/csrc/renumber.c: a numbering with no
 * locality - mean node-number span of an element above 2^20 - or BFM_RENUMBER=1)?  1: yes, and to_new[a] (if not NULL,
 * n_nodes entries) is the internal number of the caller's node a along a Morton curve through the coordinates;
 * 0: the caller's numbering is used as it is; -1: error.  Results always come back in the caller's numbering. */
int bfmx_mesh_internal_numbering(bfm_mesh_t* mesh, int32_t* to_new);

/* structured plate [0,lx]x[0,ly] with nx*ny cells: 2 triangles (a,b,d),(a,d,c) per cell, or 1 quad */
int bfmx_mesh_plate(bfm_mesh_t* mesh, bfm_state_t* state, size_t nx, size_t ny, double lx, double ly, bfm_elem_kind_t kind, bool with_edges);

/* the symbolic plan of a mesh (SELL-32 node-block pattern + element-to-nonzero map), for inspection:
 * slot (row a, position t) = slice_off[a / 32] + 32 t + a % 32; contributions are packed as
 * element << 4 | local row node << 2 | local column node, in the reference's accumulation order */
int bfmx_mesh_pattern_sizes(bfm_mesh_t* mesh, size_t* n_slices, size_t* n_slots, size_t* n_blocks, size_t* n_contributions);
int bfmx_mesh_pattern_copy(bfm_mesh_t* mesh, int32_t* slice_off, int32_t* row_len, int32_t* scol, int32_t* diag_pos, int32_t* ctr_ptr, uint32_t* ctr);

/* ---- multi-GPU: one process per GPU of one box ---------------------------------------------------------
 *
 * After bfmx_dist_init, bfm_sim_run and the bfmx_job_* stages become COLLECTIVE calls: every rank builds
 * the same simulation (same mesh, conditions, forces - SPMD, like the ranks of a torchrun job), owns a
 * contiguous block of node rows, assembles the elements touching them, and solves with the interface
 * DOFs and the CG dot products exchanged over NVLink.  Every rank ends with the complete displacement
 * field in instance->effects.  The staged bfm_system_* / bfm_matrix_* API stays single-GPU. */

#define BFMX_DIST_ID_BYTES 128

int bfmx_dist_unique_id(void* id);                       /* call on one rank, ship the bytes to the others */
int bfmx_dist_init(int rank, int world, void const* id); /* collective */
int bfmx_dist_finalize(void);
int bfmx_dist_rank(void);
int bfmx_dist_world(void);
/* "" when the per-iteration exchanges use NVLink peer memory (CUDA IPC mailboxes); otherwise why they use NCCL */
char const* bfmx_dist_peer_memory_status(void);

/* the row partition a mesh would get (host-only: works without a GPU) */
typedef struct {
	size_t first_node, end_node;   /* owned global nodes */
	size_t n_local_nodes;          /* owned + ghost */
	size_t own_begin, own_end;     /* owned nodes' local ids */
	size_t n_local_elems;
	size_t n_neighbours;
	size_t n_send;
} bfmx_partition_info_t;

int bfmx_partition_sizes(bfm_mesh_t* mesh, int rank, int world, bfmx_partition_info_t* out);
int bfmx_partition_copy(bfm_mesh_t* mesh, int rank, int world, size_t* local_to_global, size_t* local_elems, size_t* elem_to_global, int32_t* neighbours, int32_t* recv_begin, int32_t* recv_count, int32_t* send_ptr, int32_t* send_idx);

/* ---- solver internals open to inspection ------------------------------------------------------------- */

/* the coarse level the solver would build for a mesh with `target` aggregates (host-only): aggregate of
 * every node, colour of every aggregate; *n_aggregates = 0 when the mesh gets none */
int bfmx_coarse_plan(bfm_mesh_t* mesh, int target, int32_t* n_aggregates, int32_t* n_colors, int32_t* node_aggregate, int32_t* aggregate_color);

/* the aggregation hierarchy the solver's multilevel preconditioner would use for a mesh (host-only, tests):
 * level 0 = the mesh nodes, every further level the aggregates of the one below, the last one solved densely;
 * n_levels = 0 when the mesh gets none (too small, or it does not coarsen) */
#define BFMX_HIER_MAX_LEVELS 12

typedef struct {
	int32_t n_levels;
	int32_t n_nodes[BFMX_HIER_MAX_LEVELS];
	int64_t n_slots[BFMX_HIER_MAX_LEVELS];   /* padded SELL-32 slots of the level's operator */
	int64_t n_entries[BFMX_HIER_MAX_LEVELS]; /* entries of the prolongator towards the next level (0 on the last): one per node with the tentative prolongator, one per (node, neighbouring aggregate) with smoothed aggregation */
	int32_t smoothed;                        /* 1: smoothed aggregation (BFM_MG_SMOOTH, the default), V-cycle; 0: plain aggregation, W-cycle */
	int32_t pad_;
} bfmx_hier_info_t;

int bfmx_hier_info(bfm_mesh_t* mesh, bfmx_hier_info_t* info);
/* per level (any pointer may be NULL): aggregate[n] of every node (-1: left out), geometry[2 n] of every node
 * relative to its aggregate's reference point (neither on the last level); pattern_rowptr[n + 1] /
 * pattern_col: the level's operator pattern as CSR with ascending columns (call with pattern_col = NULL first
 * to learn the length) */
int bfmx_hier_level(bfm_mesh_t* mesh, int level, int32_t* aggregate, float* geometry, int32_t* pattern_rowptr, int32_t* pattern_col);

/* ---- matrices ---------------------------------------------------------------------------------------- */

/* CSR-kind matrix from a scalar CSR triple (n even: DOFs are paired into nodes); duplicates are summed.
 * Host-only until the first solve, which uploads it. */
int bfmx_matrix_csr_create(bfm_matrix_t* matrix, bfm_state_t* state, size_t n, size_t const* rowptr, size_t const* col, double const* val);

/* scalar CSR copy of a CSR-kind matrix: structural pattern (numeric zeros included), ascending columns,
 * original (un-renumbered) DOF order.  Call with NULL arrays first to learn nnz. */
int bfmx_matrix_csr_export(bfm_matrix_t* matrix, size_t* nnz, size_t* rowptr, size_t* col, double* val);

#ifdef __cplusplus
}
#endif

#endif
