/*
 * internal.h - declarations shared by the host-side C sources of libbfm (B200 build).
 * Nothing here is exported; the public surface is include/bfm/libbfm.h + include/bfm_b200.h.
 */
#ifndef BFM_INTERNAL_H
#define BFM_INTERNAL_H

#include <stdarg.h>
#include <stdint.h>

#include <bfm/libbfm.h>
#include <bfm_b200.h>

#include "gpu.h"

#define BFMI_HIDDEN __attribute__((visibility("hidden")))

/* fill state->err, print it on stderr (unless BFM_QUIET is set) and return -1 */
BFMI_HIDDEN int bfmi_fail(bfm_state_t* state, char const* file, char const* func, size_t line, char const* fmt, ...);
#define BFMI_FAIL(state, ...) bfmi_fail((state), __FILE__, __func__, __LINE__, __VA_ARGS__)

/* ---------------------------------------------------------------------------------------------
 * symbolic plan: everything that depends only on the mesh connectivity (built once per mesh)
 *
 * The global matrix is stored by NODE blocks (2x2 doubles per pair of coupled nodes) in a
 * sliced-ELL layout of 32-row slices ("SELL-32"): slice s holds block rows 32s..32s+31, padded to
 * the longest of them; slot (row a, position t) lives at  slice_off[a / 32] + 32 t + a % 32.  Both
 * the assembly and the SpMV kernels walk a slice with one warp, lane = row, so every access to the
 * value/column arrays is a contiguous 32-wide segment.
 * ------------------------------------------------------------------------------------------- */

typedef struct bfmi_plan {
	int refs;

	/* cache key (plans built from a mesh only) */
	bfm_mesh_t const* mesh;
	size_t n_nodes, n_elems;
	int kind;
	uint64_t elems_hash;

	/* host copies */
	int32_t nb;         /* block rows = nodes */
	int32_t n_slices;   /* ceil(nb / 32) */
	int64_t n_slots;    /* padded block count (multiple of 32) */
	int64_t n_blocks;   /* real blocks */
	int64_t n_ctr;      /* element contributions over all blocks */

	int32_t* slice_off; /* [n_slices + 1] first slot of each slice */
	int32_t* row_len;   /* [nb] real blocks per row */
	int32_t* scol;      /* [n_slots] block column; padding slots point at their own row */
	int32_t* diag_pos;  /* [nb] slot of the diagonal block */
	int32_t* ctr_ptr;   /* [n_slots + 1] contributor list bounds (NULL without a mesh) */
	uint32_t* ctr;      /* [n_ctr] element << 4 | local row node << 2 | local column node */
	int32_t* elems32;   /* [n_elems * kind] connectivity narrowed to 32 bits */

	bfmg_pattern_t dev; /* device arrays: built there by symbolic.cu, or mirrored by bfmi_plan_upload */
	bool on_device;
	bool built_on_device;         /* symbolic.cu built it: ctr_ptr, ctr, elems32 exist on the device only */
	size_t h2d_bytes, d2h_bytes;  /* what building / mirroring it moved over PCIe */
	bool accounted;               /* ... and whether a job has put that (and the build time) into its stats yet */
} bfmi_plan_t;

BFMI_HIDDEN bfmi_plan_t* bfmi_plan_for_mesh(bfm_state_t* state, bfm_mesh_t const* mesh); /* cached, retained */
BFMI_HIDDEN bfmi_plan_t* bfmi_plan_for_mesh_hashed(bfm_state_t* state, bfm_mesh_t const* mesh, uint64_t hash); /* hash = bfmi_mesh_hash(mesh) */
BFMI_HIDDEN bfmi_plan_t* bfmi_plan_from_csr(bfm_state_t* state, size_t n, size_t const* rowptr, size_t const* col);
BFMI_HIDDEN void bfmi_plan_retain(bfmi_plan_t* plan);
BFMI_HIDDEN void bfmi_plan_release(bfmi_plan_t* plan);
BFMI_HIDDEN void bfmi_plan_forget(bfm_mesh_t const* mesh); /* mesh is going away */
BFMI_HIDDEN int bfmi_plan_upload(bfm_state_t* state, bfmi_plan_t* plan);

/* slot of block (a, b), or -1 */
BFMI_HIDDEN int64_t bfmi_plan_find(bfmi_plan_t const* plan, int32_t a, int32_t b);

/* ---------------------------------------------------------------------------------------------
 * internal node numbering for meshes numbered without locality (renumber.c)
 * ------------------------------------------------------------------------------------------- */

typedef struct bfmi_renum {
	int refs;

	/* cache key */
	bfm_mesh_t const* orig;
	size_t n_nodes, n_elems;
	int kind;
	uint64_t hash;      /* of the caller's connectivity */

	int32_t* to_new;    /* [n_nodes] the caller's node -> internal node */
	int32_t* to_old;    /* ... and back */
	int32_t* d_to_new;  /* device copy (bfmi_renum_upload) */

	bfm_mesh_t mesh;    /* the internal mesh: same elements in the same order, nodes renumbered; no edges, no domains */
	uint64_t mesh_hash; /* of its connectivity */
} bfmi_renum_t;

/* NULL: the caller's numbering is kept.  Otherwise retained, with the coordinates of the internal mesh refreshed
 * from the caller's.  hash = bfmi_mesh_hash(mesh) */
BFMI_HIDDEN bfmi_renum_t* bfmi_renum_for_mesh(bfm_mesh_t const* mesh, uint64_t hash);
BFMI_HIDDEN void bfmi_renum_release(bfmi_renum_t* r);
BFMI_HIDDEN void bfmi_renum_forget(bfm_mesh_t const* mesh);
BFMI_HIDDEN int bfmi_renum_upload(bfmi_renum_t* r);

/* ---------------------------------------------------------------------------------------------
 * row partition of a mesh over the ranks of a multi-GPU job (partition.c)
 * ------------------------------------------------------------------------------------------- */

typedef struct bfmi_part {
	int refs;

	bfm_mesh_t const* global;
	size_t n_nodes, n_elems;
	uint64_t hash;
	int rank, world;

	size_t lo, hi;        /* owned global nodes [lo, hi) */
	int32_t n_local;      /* owned + ghost nodes */
	int32_t own_begin;    /* local ids [own_begin, own_end) are the owned nodes, in global order */
	int32_t own_end;
	size_t* l2g;          /* [n_local] local -> global node, ascending */
	size_t* elem_l2g;     /* [local.n_elems] local -> global element, ascending */

	bfm_mesh_t local;     /* coordinates + connectivity in local ids (no edges, no domains) */

	int32_t n_nbr;
	int32_t* nbr;         /* neighbour ranks, ascending */
	int32_t* recv_begin;  /* per neighbour: its ghosts are local ids [recv_begin, recv_begin + recv_count) */
	int32_t* recv_count;
	int32_t* send_ptr;    /* [n_nbr + 1] */
	int32_t n_send;
	int32_t* send_idx;    /* [n_send] owned local ids each neighbour ghosts, ascending per neighbour */

	int32_t* d_send_idx;  /* device mirror (uploaded by the job) */
} bfmi_part_t;

BFMI_HIDDEN size_t bfmi_part_first_node(size_t n_nodes, int world, int rank);
BFMI_HIDDEN int bfmi_part_owner(size_t n_nodes, int world, size_t node);
BFMI_HIDDEN int32_t bfmi_part_local(bfmi_part_t const* part, size_t global_node); /* -1 when not local */
BFMI_HIDDEN bfmi_part_t* bfmi_part_build(bfm_state_t* state, bfm_mesh_t const* mesh, int rank, int world);
BFMI_HIDDEN bfmi_part_t* bfmi_part_for_mesh(bfm_state_t* state, bfm_mesh_t const* mesh, uint64_t hash, int rank, int world); /* cached, retained */
BFMI_HIDDEN void bfmi_part_release(bfmi_part_t* part);
BFMI_HIDDEN void bfmi_part_forget(bfm_mesh_t const* mesh);
BFMI_HIDDEN uint64_t bfmi_mesh_hash(bfm_mesh_t const* mesh);
BFMI_HIDDEN uint64_t bfmi_mesh_hash_part(bfm_mesh_t const* mesh, int rank, int world); /* XOR over the ranks = bfmi_mesh_hash */

/* ---------------------------------------------------------------------------------------------
 * coarse level of the solver (coarse.c): node aggregates, their colouring, device mirrors
 * ------------------------------------------------------------------------------------------- */

typedef struct bfmi_coarse {
	int refs;

	/* cache key */
	bfm_mesh_t const* gmesh;
	size_t key_nodes, key_elems;
	uint64_t elems_hash, coords_hash;
	int rank, world;
	int32_t target;
	bool on_device;

	int32_t n_agg;
	int32_t n_colors;
	int32_t n_local;     /* nodes of the mesh this rank assembles (owned + ghost) */
	int32_t agg_span;    /* largest |g - h| over adjacent aggregates g, h */

	int32_t* agg;        /* [n_local] */
	double* wgeom;       /* [n_local][2] */
	int32_t* agg_ptr;    /* [n_agg + 1] over OWNED nodes */
	int32_t* agg_nodes;
	int32_t* color;      /* [n_agg] */
	int32_t* color_nbr;  /* [n_agg][n_colors] */

	bfmg_coarse_t dev;
} bfmi_coarse_t;

/* NULL when the mesh is too small or too degenerate for a coarse level (not an error) */
BFMI_HIDDEN bfmi_coarse_t* bfmi_coarse_build(bfm_state_t* state, bfm_mesh_t const* gmesh, bfmi_part_t const* part, int32_t target_aggregates);
BFMI_HIDDEN int bfmi_coarse_upload(bfmi_coarse_t* coarse);
BFMI_HIDDEN void bfmi_coarse_free(bfmi_coarse_t* coarse);
/* cached (one entry, keyed by mesh identity + connectivity and coordinate hashes + partition), retained;
 * *none is set when the mesh simply gets no coarse level */
BFMI_HIDDEN bfmi_coarse_t* bfmi_coarse_for_mesh(bfm_state_t* state, bfm_mesh_t const* gmesh, uint64_t elems_hash, bfmi_part_t const* part, int32_t target, bool* none);
BFMI_HIDDEN void bfmi_coarse_release(bfmi_coarse_t* coarse);
BFMI_HIDDEN void bfmi_coarse_forget(bfm_mesh_t const* gmesh);

/* ---------------------------------------------------------------------------------------------
 * aggregation hierarchy of the multilevel preconditioner (hier.c), kernels in mg.cuh
 * ------------------------------------------------------------------------------------------- */

#define BFMI_MG_SMOOTH_DEFAULT 1 /* BFM_MG_SMOOTH: 1 = smoothed aggregation on every level (V-cycle), 0 = plain aggregation (W-cycle) */

typedef struct bfmi_hier_level {
	int32_t n;            /* nodes of this level (level 0: the nodes the plan covers) */
	int32_t dofs;         /* unknowns per node: 2 on level 0, 3 above */

	/* SELL-32 node pattern of the level's operator (level 0: the plan's arrays, borrowed) */
	int32_t n_slices;
	int64_t n_slots;
	int32_t* slice_off;
	int32_t* row_len;
	int32_t* scol;
	int32_t* diag_pos;
	double* pos;          /* [n][2] reference points (level 0: the mesh coordinates, borrowed) */
	bool owns_pattern;

	/* several GPUs: the nodes [row_lo, row_hi) are this rank's (it has their pattern rows and aggregates them);
	 * a distributed level also holds ghosts, refreshed by a halo exchange before every product with its operator;
	 * a replicated level is held in full by every rank.  One GPU: row_lo = 0, row_hi = n. */
	int32_t row_lo, row_hi;
	bool distributed;
	int32_t n_nbr;
	int32_t nbr[BFMG_DIST_MAX_RANKS];        /* neighbour ranks, ascending */
	int32_t recv_begin[BFMG_DIST_MAX_RANKS]; /* per neighbour: its ghosts are the nodes [recv_begin, recv_begin + recv_count) */
	int32_t recv_count[BFMG_DIST_MAX_RANKS];
	int32_t send_ptr[BFMG_DIST_MAX_RANKS + 1];
	int32_t n_send;
	int32_t* send_idx;                       /* [n_send] owned nodes each neighbour ghosts, in the order it expects them */
	int32_t gather_first, gather_count;      /* distributed level below a replicated one: global ids of this rank's aggregates */

	/* towards the next level (absent on the last) */
	int32_t n_coarse;     /* nodes of the next level = aggregates of this one */
	int32_t n_p;          /* entries of the prolongator */
	int32_t* agg;         /* [n] aggregate of every node, -1: left out of the coarse space */
	float* geom;          /* [n][2] node position relative to its aggregate's reference point */
	int32_t* p_ptr;       /* [n + 1] prolongator by fine node */
	int32_t* p_col;       /* [n_p] coarse node of every entry */
	int32_t* r_ptr;       /* [n_coarse + 1] its transpose by coarse node */
	int32_t* r_ent;       /* [n_p] entry index */
	int32_t* r_node;      /* [n_p] fine node of that entry, ascending per coarse node */
	bool smoothed;        /* smoothed aggregation: nodes whose row is all this rank's reach their neighbours' aggregates (hier.c: build_transfer) */

	bfmg_mg_level_t dev;
} bfmi_hier_level_t;

typedef struct bfmi_hier {
	int refs;

	/* cache key */
	struct bfmi_plan const* key_plan;
	size_t key_nodes;
	uint64_t elems_hash, coords_hash, settings;
	int rank, world;
	bool on_device;

	int n_levels;         /* the last level is solved with a dense inverse */
	int32_t dense_span;   /* largest |I - J| over coupled nodes of the last level */
	bfmi_hier_level_t level[BFMG_MG_MAX_LEVELS];

	bfmi_plan_t* plan;    /* retained: level 0 borrows its pattern */
	bfmg_mg_t dev;
} bfmi_hier_t;

/* NULL when the mesh is too small or does not coarsen (not an error: the solver falls back) */
BFMI_HIDDEN bfmi_hier_t* bfmi_hier_build(bfmi_plan_t const* plan, double const* coords, bfmi_part_t const* part);
BFMI_HIDDEN int bfmi_hier_upload(bfmi_hier_t* hier, bfmg_pattern_t const* pat0, size_t* h2d_bytes);
BFMI_HIDDEN void bfmi_hier_free(bfmi_hier_t* hier);
/* cached (one entry, keyed by plan identity + coordinate hash + partition + settings), retained */
BFMI_HIDDEN bfmi_hier_t* bfmi_hier_for_plan(bfmi_plan_t* plan, double const* coords, bfmi_part_t const* part); /* collective on several GPUs */
BFMI_HIDDEN void bfmi_hier_release(bfmi_hier_t* hier);
BFMI_HIDDEN void bfmi_hier_forget(bfmi_plan_t const* plan);

/* ---------------------------------------------------------------------------------------------
 * BFM_MATRIX_KIND_CSR implementation object (matrix->csr.impl)
 * ------------------------------------------------------------------------------------------- */

typedef struct bfmi_csr {
	bfm_state_t* state;
	bfmi_plan_t* plan;

	double* d_val;      /* device, 4 * n_slots doubles: plane of (a00,a01) pairs, then plane of (a10,a11) */
	bool d_valid;       /* device copy is current */
	double* h_val;      /* host mirror, same layout (lazy) */
	bool h_valid;

	size_t* perm;       /* logical renumbering applied by bfm_perm_perm_matrix, or NULL */
	size_t* inv_perm;

	bfmx_stats_t stats; /* of the last solve */
} bfmi_csr_t;

BFMI_HIDDEN int bfmi_csr_wrap(bfm_matrix_t* matrix, bfm_state_t* state, bfmi_plan_t* plan, double* d_val);
BFMI_HIDDEN int bfmi_csr_destroy(bfm_matrix_t* matrix);
BFMI_HIDDEN int bfmi_csr_mirror(bfmi_csr_t* csr); /* make h_val valid */
BFMI_HIDDEN double bfmi_csr_get(bfm_matrix_t* matrix, size_t i, size_t j);
BFMI_HIDDEN int bfmi_csr_put(bfm_matrix_t* matrix, size_t i, size_t j, double val, bool add);
BFMI_HIDDEN size_t bfmi_csr_bandwidth(bfm_matrix_t* matrix);
BFMI_HIDDEN int bfmi_csr_solve(bfm_matrix_t* matrix, bfm_vec_t* y);
BFMI_HIDDEN int bfmi_csr_copy(bfm_matrix_t* dst, bfm_matrix_t* src);
BFMI_HIDDEN int bfmi_csr_rcm(bfm_perm_t* perm, bfm_matrix_t* matrix);
BFMI_HIDDEN int bfmi_csr_set_perm(bfm_matrix_t* matrix, size_t const* perm, size_t n);

/* RCM driver shared by the dense and the sparse front ends (perm.c): row_nnz(i, out, ctx) writes
 * the ascending column indices of the numerically non-zero entries of row i and returns their count */
typedef size_t (*bfmi_row_fn_t)(size_t i, size_t* out, void* ctx);
BFMI_HIDDEN int bfmi_rcm(bfm_perm_t* perm, size_t n, size_t max_row, bfmi_row_fn_t row_nnz, void* ctx);

/* solver options resolved from the environment (BFM_CG_TOL, BFM_CG_MAXIT, ...) */
BFMI_HIDDEN void bfmi_pcg_options(size_t n, bfmg_pcg_opts_t* opts);

#endif
