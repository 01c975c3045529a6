/*
 * hier.c - host side of the solver's multilevel preconditioner: the aggregation hierarchy.
 *
 * The reference solves with a direct band LU (matrix.c:253-404); this build solves with conjugate
 * gradients preconditioned by an aggregation multigrid cycle (mg.cuh explains the numerics).  What is
 * decided here depends only on the mesh (connectivity + coordinates) and is built once per mesh:
 *
 *   level 0   the mesh's nodes (2 unknowns each); its operator is the assembled matrix (plan.c's SELL-32
 *             node-block pattern)
 *   level l   the aggregates of level l - 1 (3 unknowns each: the rigid-body modes of plane elasticity - two
 *             translations and the rotation about the aggregate's reference point)
 *   last      few enough nodes for a dense inverse (coarse.cuh's Gauss-Jordan)
 *
 * Aggregates are the connected pieces of the cells of a uniform grid of bins over the level's bounding box
 * (pieces too small to carry three independent modes join a neighbour).  Between two levels sits the
 * prolongator P, stored by its sparsity structure (CSR by fine node + its transpose by coarse node); its
 * values are computed on the device for every solve because they carry the matrix's diagonal scaling.  With
 * smoothed aggregation (BFM_MG_SMOOTH, the default) P = (I - w A^) P~: the row of a node has an entry for the
 * aggregate of every neighbour (build_transfer below), with the tentative prolongator P~ one entry, its own
 * aggregate's.  The next level's operator P^T A P is computed on the device too (k_mg_ap + k_mg_ptq, or
 * k_mg_rap); its sparsity pattern is computed here, symbolically, as the same product.
 *
 * Several GPUs (one rank per GPU, partition.c): every rank builds the hierarchy of ITS rows.  Aggregates are formed
 * from owned nodes only, so they never straddle two ranks - and only nodes with no ghost among their columns are
 * smoothed, so no prolongator row does either: restriction and prolongation stay local and only the
 * products with a level's operator need a halo exchange, exactly as on the mesh level.  A level is DISTRIBUTED
 * (local numbering: owned nodes first, then the ghosts of each neighbour, ascending by the owner's numbering)
 * while it is large; from the first level of at most BFM_MG_REPLICATED_NODES nodes on it is REPLICATED: every
 * rank holds all of it (global numbering, rank after rank) and runs it redundantly - cheaper than exchanging
 * halos of a few hundred nodes several times per cycle.  What the ranks must tell each other while building -
 * aggregate counts, the aggregates of ghost nodes, the rows of the first replicated level - goes through the
 * communicator of dist.cu (a handful of small collectives, once per mesh).
 *
 * Everything is deterministic: ids follow the smallest member, sets are sorted.
 */
#include "internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include <stdio.h>
#include <time.h>

#define SLICE 32

/* BFM_MG_VERBOSE=1: where the time of the hierarchy build goes, on stderr */
static double hier_clock(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec * 1e3 + (double) ts.tv_nsec * 1e-6;
}

#define MIN_NODES_LEVEL0 3 /* nodes an aggregate of mesh nodes needs for three independent modes */

static inline int64_t slot_of(int32_t const* slice_off, int32_t a, int32_t t) {
	return (int64_t) slice_off[a / SLICE] + (int64_t) t * SLICE + a % SLICE;
}

static void level_free(bfmi_hier_level_t* L) {
	if (L->owns_pattern) {
		free(L->slice_off);
		free(L->row_len);
		free(L->scol);
		free(L->diag_pos);
		free(L->pos);
	}

	free(L->send_idx);
	bfmg_free(L->dev.send_idx);

	free(L->agg);
	free(L->geom);
	free(L->p_ptr);
	free(L->p_col);
	free(L->r_ptr);
	free(L->r_ent);
	free(L->r_node);

	bfmg_mg_level_t* const d = &L->dev;

	if (L->owns_pattern) {
		bfmg_free(d->slice_off);
		bfmg_free(d->scol);
		bfmg_free(d->diag_pos);
		bfmg_free(d->row_len);
	}

	bfmg_free(d->agg);
	bfmg_free(d->geom);
	bfmg_free(d->p_ptr);
	bfmg_free(d->p_col);
	bfmg_free(d->r_ptr);
	bfmg_free(d->r_ent);
	bfmg_free(d->r_node);
}

void bfmi_hier_free(bfmi_hier_t* h) {
	if (h == NULL) {
		return;
	}

	for (int l = 0; l < h->n_levels; l++) {
		level_free(&h->level[l]);
	}

	bfmi_plan_release(h->plan);
	free(h);
}

/* ---- aggregation of one level --------------------------------------------------------------------------
 *
 * agg[a] = aggregate of node a, compacted to 0 .. n_agg - 1 in order of each aggregate's smallest node, or -1 for
 * nodes left out of the coarse space (isolated nodes: nothing to interpolate from).  Returns n_agg (0: give up). */

/* root of v without touching the forest: safe while other threads read it too */
static inline int32_t find_root(int32_t const* parent, int32_t v) {
	while (parent[v] != v) {
		v = parent[v];
	}

	return v;
}

static int32_t aggregate_level(bfmi_hier_level_t const* L, int64_t target, int32_t min_nodes, int32_t* agg) {
	int32_t const n = L->n;
	double const* const pos = L->pos;

	int32_t const lo = L->row_lo, hi = L->row_hi; /* the nodes this rank aggregates: all of them on one GPU */
	bool const big = hi - lo > 100000;            /* worth the host threads */

	double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
	bool bad = false;

#pragma omp parallel for schedule(static) if (n > 100000)
	for (int32_t a = 0; a < n; a++) {
		agg[a] = -1;
	}

#pragma omp parallel for schedule(static) reduction(min : x0, y0) reduction(max : x1, y1) reduction(|| : bad) if (big)
	for (int32_t a = lo; a < hi; a++) {
		double const x = pos[2 * (size_t) a + 0];
		double const y = pos[2 * (size_t) a + 1];

		bad = bad || !(x == x) || !(y == y) || isinf(x) || isinf(y);

		x0 = x < x0 ? x : x0, x1 = x > x1 ? x : x1;
		y0 = y < y0 ? y : y0, y1 = y > y1 ? y : y1;
	}

	if (bad) {
		return 0;
	}

	double wx = x1 - x0;
	double wy = y1 - y0;

	if (!(wx > 0) && !(wy > 0)) {
		return 0; /* all nodes in one point: nothing to bin */
	}

	/* a degenerate direction (nodes on a line) gets one row of bins */
	double const thin = 1e-12 * (wx > wy ? wx : wy);

	wx = wx > thin ? wx : thin;
	wy = wy > thin ? wy : thin;

	target = target < 1 ? 1 : target;

	double const side = sqrt(wx * wy / (double) target);
	int64_t nbx = (int64_t) floor(wx / side + 0.5);
	int64_t nby = (int64_t) floor(wy / side + 0.5);

	nbx = nbx < 1 ? 1 : nbx;
	nby = nby < 1 ? 1 : nby;

	while (nbx * nby > 4 * target && (nbx > 1 || nby > 1)) { /* extreme aspect ratios */
		nbx > nby ? nbx-- : nby--;
	}

	int32_t* const bin = malloc(((size_t) n + 1) * sizeof *bin);
	int32_t* const parent = malloc(((size_t) n + 1) * sizeof *parent);
	int32_t* const size = calloc((size_t) n + 1, sizeof *size);

	if (bin == NULL || parent == NULL || size == NULL) {
		free(bin);
		free(parent);
		free(size);
		return 0;
	}

#pragma omp parallel for schedule(static) if (n > 100000)
	for (int32_t a = 0; a < n; a++) {
		bin[a] = -1;
		parent[a] = a;
	}

#pragma omp parallel for schedule(static) if (big)
	for (int32_t a = lo; a < hi; a++) {
		int64_t bx = (int64_t) ((pos[2 * (size_t) a + 0] - x0) / wx * (double) nbx);
		int64_t by = (int64_t) ((pos[2 * (size_t) a + 1] - y0) / wy * (double) nby);

		bx = bx >= nbx ? nbx - 1 : (bx < 0 ? 0 : bx);
		by = by >= nby ? nby - 1 : (by < 0 ? 0 : by);

		bin[a] = (int32_t) (by * nbx + bx);
	}

#define FIND(v, out)                               \
	do {                                           \
		int32_t r_ = (v);                          \
		while (parent[r_] != r_) {                 \
			parent[r_] = parent[parent[r_]];       \
			r_ = parent[r_];                       \
		}                                          \
		(out) = r_;                                \
	} while (0)

	/* connected pieces of every bin: nodes of one bin joined through couplings of the operator.  The smaller
	 * root wins, so the result does not depend on traversal order; on a solid mesh a bin is one piece, on
	 * truss-like geometry a bin can cut through members that do not touch and each becomes its own aggregate.
	 * Unions never leave a bin, so the bins are independent: nodes are bucketed by bin (counted and dropped with
	 * integer atomics, then every bucket put in ascending order by the thread that owns it - the same buckets a
	 * serial counting sort fills) and the buckets shared out over the host threads. */

	{
		int64_t const n_bins = nbx * nby;
		int64_t* const start = calloc((size_t) n_bins + 2, sizeof *start);
		int32_t* const order = malloc(((size_t) (hi - lo) + 1) * sizeof *order);

		if (start == NULL || order == NULL) {
			free(start);
			free(order);
			free(bin);
			free(parent);
			free(size);
			return 0;
		}

#pragma omp parallel for schedule(static) if (big)
		for (int32_t a = lo; a < hi; a++) {
			__atomic_fetch_add(&start[bin[a] + 2], 1, __ATOMIC_RELAXED);
		}

		for (int64_t b = 0; b < n_bins; b++) {
			start[b + 2] += start[b + 1];
		}

#pragma omp parallel for schedule(static) if (big)
		for (int32_t a = lo; a < hi; a++) {
			order[__atomic_fetch_add(&start[bin[a] + 1], 1, __ATOMIC_RELAXED)] = a;
		}

#pragma omp parallel for schedule(dynamic, 64) if (big)
		for (int64_t b = 0; b < n_bins; b++) {
			for (int64_t i = start[b] + 1; i < start[b + 1]; i++) { /* ascending nodes inside the bin */
				int32_t const cur = order[i];
				int64_t j = i;

				for (; j > start[b] && order[j - 1] > cur; j--) {
					order[j] = order[j - 1];
				}

				order[j] = cur;
			}

			for (int64_t i = start[b]; i < start[b + 1]; i++) {
				int32_t const a = order[i];

				for (int32_t t = 0; t < L->row_len[a]; t++) {
					int32_t const c = L->scol[slot_of(L->slice_off, a, t)];

					if (c < a && c >= lo && bin[c] == bin[a]) {
						int32_t ra, rc;

						FIND(a, ra);
						FIND(c, rc);

						if (ra != rc) {
							if (ra < rc) {
								parent[rc] = ra;
							}

							else {
								parent[ra] = rc;
							}
						}
					}
				}
			}
		}

		free(start);
		free(order);
	}

	/* pieces below min_nodes join a piece they are coupled to (a few passes: a chain of tiny pieces needs more
	 * than one).  The joining itself goes node by node in ascending order - its outcome depends on that order - but
	 * only nodes of pieces that are too small when a pass starts can take part (pieces only grow), and on a solid mesh
	 * there are none: they are found in parallel first */

	int32_t* small = NULL;

	for (int pass = 0; pass < 4 && min_nodes > 1; pass++) {
		memset(size, 0, ((size_t) n + 1) * sizeof *size);

#pragma omp parallel for schedule(static) if (big)
		for (int32_t a = lo; a < hi; a++) {
			__atomic_fetch_add(&size[find_root(parent, a)], 1, __ATOMIC_RELAXED);
		}

		int64_t n_small = 0;

#pragma omp parallel for schedule(static) reduction(+ : n_small) if (big)
		for (int32_t a = lo; a < hi; a++) {
			n_small += size[find_root(parent, a)] < min_nodes;
		}

		if (n_small == 0) {
			break;
		}

		free(small);
		small = malloc(((size_t) n_small + 1) * sizeof *small);

		if (small == NULL) {
			free(bin);
			free(parent);
			free(size);
			return 0;
		}

		n_small = 0;

		for (int32_t a = lo; a < hi; a++) {
			if (size[find_root(parent, a)] < min_nodes) {
				small[n_small++] = a;
			}
		}

		bool changed = false;

		for (int64_t i = 0; i < n_small; i++) {
			int32_t const a = small[i];
			int32_t ra;
			FIND(a, ra);

			if (size[ra] >= min_nodes) {
				continue;
			}

			for (int32_t t = 0; t < L->row_len[a]; t++) {
				int32_t const b = L->scol[slot_of(L->slice_off, a, t)];
				int32_t rb;

				if (b == a || b < lo || b >= hi) {
					continue;
				}

				FIND(b, rb);

				if (rb != ra) {
					int32_t const total = size[ra] + size[rb];
					int32_t const root = ra < rb ? ra : rb;

					parent[ra] = root;
					parent[rb] = root;
					size[root] = total;
					changed = true;
					break;
				}
			}
		}

		if (!changed) {
			break;
		}
	}

	free(small);

	memset(size, 0, ((size_t) n + 1) * sizeof *size);

#pragma omp parallel for schedule(static) if (big)
	for (int32_t a = lo; a < hi; a++) {
		__atomic_fetch_add(&size[find_root(parent, a)], 1, __ATOMIC_RELAXED);
	}

#undef FIND

	/* compact: ids in order of each piece's smallest node (its root); pieces still too small are left out.
	 * bin[] is free now: reused as root -> id (a chunked prefix sum over the roots that are kept) */

	int32_t n_agg = 0;

	{
		int64_t const chunk = 1 << 16;
		int64_t const n_chunks = ((int64_t) (hi - lo) + chunk - 1) / chunk;
		int32_t* const first_id = calloc((size_t) n_chunks + 1, sizeof *first_id);

		if (first_id == NULL) {
			free(bin);
			free(parent);
			free(size);
			return 0;
		}

#pragma omp parallel for schedule(static) if (big)
		for (int64_t c = 0; c < n_chunks; c++) {
			int32_t const end = lo + (c + 1) * chunk < hi ? (int32_t) (lo + (c + 1) * chunk) : hi;
			int32_t kept = 0;

			for (int32_t a = (int32_t) (lo + c * chunk); a < end; a++) {
				kept += parent[a] == a && size[a] >= min_nodes;
			}

			first_id[c + 1] = kept;
		}

		for (int64_t c = 0; c < n_chunks; c++) {
			first_id[c + 1] += first_id[c];
		}

		n_agg = first_id[n_chunks];

#pragma omp parallel for schedule(static) if (big)
		for (int64_t c = 0; c < n_chunks; c++) {
			int32_t const end = lo + (c + 1) * chunk < hi ? (int32_t) (lo + (c + 1) * chunk) : hi;
			int32_t id = first_id[c];

			for (int32_t a = (int32_t) (lo + c * chunk); a < end; a++) {
				if (parent[a] == a) {
					bin[a] = size[a] >= min_nodes ? id++ : -1;
				}
			}
		}

		free(first_id);
	}

#pragma omp parallel for schedule(static) if (big)
	for (int32_t a = lo; a < hi; a++) {
		agg[a] = bin[find_root(parent, a)];
	}

	free(bin);
	free(parent);
	free(size);

	return n_agg;
}

/* ---- the structures between level l and level l + 1 -------------------------------------------------------- */

/* CSR of the prolongator (one entry per node that has an aggregate - ghosts included, the Galerkin product needs
 * their rows) and its transpose over the nodes this rank owns (restriction sums over owned nodes only) */
/* Smoothed aggregation (L->smoothed): the prolongator is (I - w A^) P~ instead of the tentative P~, so the row of a node
 * reaches the aggregates of all its neighbours.  Only nodes whose whole row is this rank's - owned, no ghost among
 * the columns - are smoothed: their entries name this rank's aggregates only, so restriction, prolongation and the
 * Galerkin rows of this rank's aggregates stay local, exactly as with the tentative prolongator (a rank sees the plain,
 * one-entry rows of its ghosts, which is what their owner gives them too: interface nodes are never smoothed).  On one
 * GPU and on replicated levels every node is smoothed.  mg.cuh (k_mg_smooth) applies the same rule. */
static inline bool smoothable(bfmi_hier_level_t const* L, int32_t a) {
	if (!L->smoothed || a < L->row_lo || a >= L->row_hi || L->agg[a] < 0) {
		return false;
	}

	for (int32_t t = 0; t < L->row_len[a]; t++) {
		int32_t const b = L->scol[slot_of(L->slice_off, a, t)];

		if (b < L->row_lo || b >= L->row_hi) {
			return false;
		}
	}

	return true;
}

/* the coarse nodes of the entries of node a, ascending, into out (room for row_len + 1); returns their number */
static inline int32_t entries_of(bfmi_hier_level_t const* L, int32_t a, int32_t* out) {
	if (L->agg[a] < 0) {
		return 0;
	}

	int32_t cnt = 0;

	out[cnt++] = L->agg[a];

	if (smoothable(L, a)) {
		for (int32_t t = 0; t < L->row_len[a]; t++) {
			int32_t const g = L->agg[L->scol[slot_of(L->slice_off, a, t)]];
			int32_t i = 0;

			if (g < 0) {
				continue;
			}

			for (; i < cnt && out[i] != g; i++) {
			}

			if (i == cnt) {
				out[cnt++] = g;
			}
		}

		for (int32_t i = 1; i < cnt; i++) {
			int32_t const cur = out[i];
			int32_t j = i;

			for (; j > 0 && out[j - 1] > cur; j--) {
				out[j] = out[j - 1];
			}

			out[j] = cur;
		}
	}

	return cnt;
}

static int build_transfer(bfmi_hier_level_t* L) {
	int32_t const n = L->n;
	int32_t const nc = L->n_coarse;
	bool const big = n > 100000;

	L->p_ptr = malloc(((size_t) n + 1) * sizeof *L->p_ptr);
	L->r_ptr = calloc((size_t) nc + 2, sizeof *L->r_ptr);

	if (L->p_ptr == NULL || L->r_ptr == NULL) {
		return -1;
	}

	int32_t longest = 1;

	if (L->smoothed) {
		for (int32_t a = L->row_lo; a < L->row_hi; a++) {
			longest = L->row_len[a] > longest ? L->row_len[a] : longest;
		}
	}

	/* entries per node (p_ptr[a + 1] for now), then p_ptr: their prefix sum, by chunks */

#pragma omp parallel if (big)
	{
		int32_t* const scratch = malloc(((size_t) longest + 2) * sizeof *scratch);

#pragma omp for schedule(static)
		for (int32_t a = 0; a < n; a++) {
			L->p_ptr[a + 1] = scratch != NULL ? entries_of(L, a, scratch) : -1;
		}

		free(scratch);
	}

	int64_t const chunk = 1 << 16;
	int64_t const n_chunks = ((int64_t) n + chunk - 1) / chunk;
	int64_t* const first = calloc((size_t) n_chunks + 1, sizeof *first);

	if (first == NULL) {
		return -1;
	}

	bool failed = false;

#pragma omp parallel for schedule(static) reduction(|| : failed) if (big)
	for (int64_t c = 0; c < n_chunks; c++) {
		int32_t const end = (c + 1) * chunk < n ? (int32_t) ((c + 1) * chunk) : n;
		int64_t cnt = 0;

		for (int32_t a = (int32_t) (c * chunk); a < end; a++) {
			failed = failed || L->p_ptr[a + 1] < 0;
			cnt += L->p_ptr[a + 1];
		}

		first[c + 1] = cnt;
	}

	for (int64_t c = 0; c < n_chunks; c++) {
		first[c + 1] += first[c];
	}

	if (failed || first[n_chunks] > INT32_MAX / 9) { /* nine value planes are indexed with 32-bit entries */
		free(first);
		return -1;
	}

	int32_t const n_p = (int32_t) first[n_chunks];

	L->p_ptr[0] = 0;

#pragma omp parallel for schedule(static) if (big)
	for (int64_t c = 0; c < n_chunks; c++) {
		int32_t const end = (c + 1) * chunk < n ? (int32_t) ((c + 1) * chunk) : n;
		int32_t at = (int32_t) first[c];

		for (int32_t a = (int32_t) (c * chunk); a < end; a++) {
			at += L->p_ptr[a + 1];
			L->p_ptr[a + 1] = at; /* chunk c only writes p_ptr[c * chunk + 1 .. end]: no overlap with its neighbours */
		}
	}

	free(first);

	L->n_p = n_p;

	L->p_col = malloc(((size_t) n_p + 1) * sizeof *L->p_col);
	L->r_ent = malloc(((size_t) n_p + 1) * sizeof *L->r_ent);
	L->r_node = malloc(((size_t) n_p + 1) * sizeof *L->r_node);

	if (L->p_col == NULL || L->r_ent == NULL || L->r_node == NULL) {
		return -1;
	}

#pragma omp parallel if (big)
	{
		int32_t* const scratch = malloc(((size_t) longest + 2) * sizeof *scratch);

#pragma omp for schedule(static)
		for (int32_t a = 0; a < n; a++) {
			int32_t const cnt = L->p_ptr[a + 1] - L->p_ptr[a];

			if (cnt == 0 || scratch == NULL) {
				continue;
			}

			entries_of(L, a, scratch);
			memcpy(&L->p_col[L->p_ptr[a]], scratch, (size_t) cnt * sizeof *scratch);

			if (a >= L->row_lo && a < L->row_hi) {
				for (int32_t i = 0; i < cnt; i++) {
					__atomic_fetch_add(&L->r_ptr[scratch[i] + 2], 1, __ATOMIC_RELAXED);
				}
			}
		}

		free(scratch);
	}

	for (int32_t g = 0; g < nc; g++) {
		L->r_ptr[g + 2] += L->r_ptr[g + 1];
	}

	/* every coarse node's list of (owned fine node, entry), ascending by node: dropped in with an atomic cursor, then
	 * each list put in order by one thread (a node has at most one entry per coarse node) */

#pragma omp parallel for schedule(static) if (big)
	for (int32_t a = L->row_lo; a < L->row_hi; a++) {
		for (int32_t e = L->p_ptr[a]; e < L->p_ptr[a + 1]; e++) {
			int32_t const at = __atomic_fetch_add(&L->r_ptr[L->p_col[e] + 1], 1, __ATOMIC_RELAXED);

			L->r_node[at] = a;
			L->r_ent[at] = e;
		}
	}

#pragma omp parallel for schedule(dynamic, 1024) if (big)
	for (int32_t g = 0; g < nc; g++) {
		int32_t const beg = L->r_ptr[g], end = L->r_ptr[g + 1];

		for (int32_t i = beg + 1; i < end; i++) {
			int32_t const cur = L->r_node[i];
			int32_t const cur_ent = L->r_ent[i];
			int32_t j = i;

			for (; j > beg && L->r_node[j - 1] > cur; j--) {
				L->r_node[j] = L->r_node[j - 1];
				L->r_ent[j] = L->r_ent[j - 1];
			}

			L->r_node[j] = cur;
			L->r_ent[j] = cur_ent;
		}
	}

	return 0;
}

static int cmp_i32(void const* a, void const* b) {
	int32_t const x = *(int32_t const*) a;
	int32_t const y = *(int32_t const*) b;
	return x < y ? -1 : x > y;
}

static int cmp_i64(void const* a, void const* b) {
	int64_t const x = *(int64_t const*) a;
	int64_t const y = *(int64_t const*) b;
	return x < y ? -1 : x > y;
}

/* rows [first, first + count) of the pattern of P^T A P: coarse node I couples to J when some node in the support of
 * column I of P is coupled (through A) to a node in the support of column J.  rows[i] (malloc'd, ascending) and
 * lens[i] for i = I - first. */
static int coarse_rows(bfmi_hier_level_t const* L, int32_t first, int32_t count, int32_t** rows, int32_t* lens) {
	int32_t const nc = L->n_coarse;

	int const n_threads =
#ifdef _OPENMP
		count > 20000 ? omp_get_max_threads() : 1;
#else
		1;
#endif

	int32_t* const marks = malloc((size_t) n_threads * ((size_t) nc + 1) * sizeof *marks);

	if (marks == NULL) {
		return -1;
	}

	memset(marks, 0xff, (size_t) n_threads * ((size_t) nc + 1) * sizeof *marks);

	bool failed = false;

#pragma omp parallel num_threads(n_threads)
	{
#ifdef _OPENMP
		int32_t* const mark = marks + (size_t) omp_get_thread_num() * ((size_t) nc + 1);
#else
		int32_t* const mark = marks;
#endif
		int32_t cap = 64;
		int32_t* buf = malloc((size_t) cap * sizeof *buf);

#pragma omp for schedule(dynamic, 256)
		for (int32_t i = 0; i < count; i++) {
			int32_t const I = first + i;
			int32_t cnt = 0;

			if (buf == NULL) {
				failed = true;
				continue;
			}

			mark[I] = I;
			buf[cnt++] = I; /* the diagonal always exists */

			for (int32_t at = L->r_ptr[I]; at < L->r_ptr[I + 1]; at++) {
				int32_t const a = L->r_node[at];

				for (int32_t t = 0; t < L->row_len[a]; t++) {
					int32_t const b = L->scol[slot_of(L->slice_off, a, t)];

					for (int32_t e = L->p_ptr[b]; e < L->p_ptr[b + 1]; e++) {
						int32_t const J = L->p_col[e];

						if (mark[J] != I) {
							mark[J] = I;

							if (cnt == cap) {
								int32_t* const grown = realloc(buf, (size_t) cap * 2 * sizeof *buf);

								if (grown == NULL) {
									failed = true;
									break;
								}

								buf = grown;
								cap *= 2;
							}

							buf[cnt++] = J;
						}
					}
				}
			}

			if (cnt <= 32) { /* the usual row: a handful of columns */
				for (int32_t u = 1; u < cnt; u++) {
					int32_t const cur = buf[u];
					int32_t v = u;

					for (; v > 0 && buf[v - 1] > cur; v--) {
						buf[v] = buf[v - 1];
					}

					buf[v] = cur;
				}
			}

			else {
				qsort(buf, (size_t) cnt, sizeof *buf, cmp_i32);
			}

			rows[i] = malloc((size_t) cnt * sizeof **rows);

			if (rows[i] == NULL) {
				failed = true;
				continue;
			}

			memcpy(rows[i], buf, (size_t) cnt * sizeof *buf);
			lens[i] = cnt;
		}

		free(buf);
	}

	free(marks);
	return failed ? -1 : 0;
}

/* SELL-32 layout of n_rows pattern rows (N->n vector entries: the columns may reach beyond the rows - ghosts) */
static int layout_level(bfmi_hier_level_t* N, int32_t n_rows, int32_t* const* rows, int32_t const* lens) {
	N->owns_pattern = true;
	N->row_len = calloc((size_t) n_rows + 1, sizeof *N->row_len);
	N->n_slices = (n_rows + SLICE - 1) / SLICE;
	N->slice_off = malloc(((size_t) N->n_slices + 1) * sizeof *N->slice_off);

	if (N->row_len == NULL || N->slice_off == NULL) {
		return -1;
	}

	memcpy(N->row_len, lens, (size_t) n_rows * sizeof *lens);

	int64_t off = 0;

	for (int32_t s = 0; s < N->n_slices; s++) {
		int32_t longest = 0;

		for (int32_t a = s * SLICE; a < n_rows && a < (s + 1) * SLICE; a++) {
			longest = lens[a] > longest ? lens[a] : longest;
		}

		N->slice_off[s] = (int32_t) off;
		off += (int64_t) longest * SLICE;

		if (off > INT32_MAX / 9) { /* nine value planes are indexed with 32-bit slots */
			return -1;
		}
	}

	N->slice_off[N->n_slices] = (int32_t) off;
	N->n_slots = off;

	N->scol = malloc(((size_t) N->n_slots + 1) * sizeof *N->scol);
	N->diag_pos = malloc(((size_t) n_rows + 1) * sizeof *N->diag_pos);

	if (N->scol == NULL || N->diag_pos == NULL) {
		return -1;
	}

	/* padding slots point at their own row (clamped) and carry zeros */

	for (int32_t s = 0; s < N->n_slices; s++) {
		for (int32_t slot = N->slice_off[s]; slot < N->slice_off[s + 1]; slot++) {
			int32_t const a = s * SLICE + (slot - N->slice_off[s]) % SLICE;
			N->scol[slot] = a < n_rows ? a : n_rows - 1;
		}
	}

	for (int32_t I = 0; I < n_rows; I++) {
		for (int32_t t = 0; t < lens[I]; t++) {
			int64_t const slot = slot_of(N->slice_off, I, t);

			N->scol[slot] = rows[I][t];

			if (rows[I][t] == I) {
				N->diag_pos[I] = (int32_t) slot;
			}
		}
	}

	return 0;
}

/* ---- what the ranks tell each other while building (device-staged: the communicator lives in dist.cu) ------- */

static int comm_allgather(double const* mine, size_t count, double* all) {
	int const world = bfmg_dist_world();
	double *d_send = NULL, *d_recv = NULL;
	int rv = -1;

	if (
		bfmg_alloc((void**) &d_send, (count + 1) * sizeof(double)) == 0 &&
		bfmg_alloc((void**) &d_recv, ((size_t) world * count + 1) * sizeof(double)) == 0 &&
		bfmg_upload(d_send, mine, count * sizeof(double)) == 0 &&
		bfmg_dist_allgather_f64(d_send, d_recv, (int) count) == 0 &&
		bfmg_download(all, d_recv, (size_t) world * count * sizeof(double)) == 0
	) {
		rv = 0;
	}

	bfmg_free(d_send);
	bfmg_free(d_recv);

	return rv;
}

/* every rank's variable-length list, one after the other in rank order: *out (malloc'd), counts[world] */
static int comm_allgather_v(double const* mine, size_t count, double** out, size_t* counts) {
	int const world = bfmg_dist_world();
	double const my_count = (double) count;
	double all_counts[BFMG_DIST_MAX_RANKS];

	if (comm_allgather(&my_count, 1, all_counts) < 0) {
		return -1;
	}

	size_t longest = 1, total = 0;

	for (int r = 0; r < world; r++) {
		counts[r] = (size_t) all_counts[r];
		longest = counts[r] > longest ? counts[r] : longest;
		total += counts[r];
	}

	double* const padded = calloc(longest, sizeof *padded);
	double* const all = malloc((size_t) world * longest * sizeof *all);

	*out = malloc((total + 1) * sizeof **out);

	if (padded == NULL || all == NULL || *out == NULL) {
		free(padded);
		free(all);
		free(*out);
		*out = NULL;
		return -1;
	}

	memcpy(padded, mine, count * sizeof *mine);

	int const rv = comm_allgather(padded, longest, all);

	if (rv == 0) {
		size_t at = 0;

		for (int r = 0; r < world; r++) {
			memcpy(*out + at, all + (size_t) r * longest, counts[r] * sizeof(double));
			at += counts[r];
		}
	}

	else {
		free(*out);
		*out = NULL;
	}

	free(padded);
	free(all);

	return rv;
}

/* ghost entries of vec (two doubles per node of level L) from their owners */
static int comm_halo(bfmi_hier_level_t const* L, double* vec) {
	if (L->n_nbr == 0) {
		return 0;
	}

	bfmg_halo_t halo = {0};
	double *d_vec = NULL, *d_buf = NULL;
	int32_t* d_idx = NULL;
	int rv = -1;

	halo.n_nbr = L->n_nbr;
	halo.nbr = L->nbr;
	halo.recv_begin = L->recv_begin;
	halo.recv_count = L->recv_count;
	halo.send_ptr = L->send_ptr;
	halo.n_send = L->n_send;

	if (
		bfmg_alloc((void**) &d_vec, ((size_t) L->n + 1) * 2 * sizeof(double)) == 0 &&
		bfmg_alloc((void**) &d_buf, ((size_t) L->n_send + 1) * 2 * sizeof(double)) == 0 &&
		bfmg_alloc((void**) &d_idx, ((size_t) L->n_send + 1) * sizeof(int32_t)) == 0 &&
		bfmg_upload(d_vec, vec, (size_t) L->n * 2 * sizeof(double)) == 0 &&
		(L->n_send == 0 || bfmg_upload(d_idx, L->send_idx, (size_t) L->n_send * sizeof(int32_t)) == 0)
	) {
		halo.d_send_idx = d_idx;

		if (bfmg_dist_halo(&halo, d_vec, d_buf) == 0 && bfmg_download(vec, d_vec, (size_t) L->n * 2 * sizeof(double)) == 0) {
			rv = 0;
		}
	}

	bfmg_free(d_vec);
	bfmg_free(d_buf);
	bfmg_free(d_idx);

	return rv;
}

/* ---- the hierarchy ---------------------------------------------------------------------------------------- */

static int64_t env_i64(char const* name, int64_t fallback) {
	char const* const env = getenv(name);
	return env != NULL && env[0] != 0 ? atoll(env) : fallback;
}

/* reference points of the aggregates of L (centroids over the nodes this rank aggregated) into cen[2 * n_agg], and
 * the geometry of the owned nodes relative to them (ghosts: filled by the exchange, or unused) */
static int centroids(bfmi_hier_level_t* L, int32_t n_agg, int32_t id_shift, double* cen) {
	int32_t* const count = calloc((size_t) n_agg + 1, sizeof *count);

	L->geom = calloc(((size_t) L->n + 1) * 2, sizeof *L->geom);

	if (count == NULL || L->geom == NULL) {
		free(count);
		return -1;
	}

	if (L->r_ptr != NULL && id_shift == 0) {
		/* the transfer lists exist (build_transfer): every aggregate's owned nodes in ascending order - the order the
		 * loop below adds them in, so the sums are the same to the last bit, one aggregate per thread */

#pragma omp parallel for schedule(static) if (n_agg > 20000)
		for (int32_t g = 0; g < n_agg; g++) {
			double sx = 0, sy = 0;

			int32_t members = 0;

			for (int32_t at = L->r_ptr[g]; at < L->r_ptr[g + 1]; at++) {
				if (L->agg[L->r_node[at]] == g) { /* a smoothed prolongator lists the neighbours of the aggregate too */
					sx += L->pos[2 * (size_t) L->r_node[at] + 0];
					sy += L->pos[2 * (size_t) L->r_node[at] + 1];
					members++;
				}
			}

			cen[2 * (size_t) g + 0] = sx;
			cen[2 * (size_t) g + 1] = sy;
			count[g] = members;
		}
	}

	else {
		for (int32_t a = L->row_lo; a < L->row_hi; a++) { /* fixed order: identical wherever it is computed */
			int32_t const g = L->agg[a] - id_shift;

			if (L->agg[a] >= 0) {
				cen[2 * (size_t) g + 0] += L->pos[2 * (size_t) a + 0];
				cen[2 * (size_t) g + 1] += L->pos[2 * (size_t) a + 1];
				count[g]++;
			}
		}
	}

	for (int32_t g = 0; g < n_agg; g++) {
		cen[2 * (size_t) g + 0] /= (double) count[g];
		cen[2 * (size_t) g + 1] /= (double) count[g];
	}

	free(count);

#pragma omp parallel for schedule(static) if (L->row_hi - L->row_lo > 100000)
	for (int32_t a = L->row_lo; a < L->row_hi; a++) {
		int32_t const g = L->agg[a] - id_shift;

		if (L->agg[a] >= 0) {
			L->geom[2 * (size_t) a + 0] = (float) (L->pos[2 * (size_t) a + 0] - cen[2 * (size_t) g + 0]);
			L->geom[2 * (size_t) a + 1] = (float) (L->pos[2 * (size_t) a + 1] - cen[2 * (size_t) g + 1]);
		}
	}

	return 0;
}

/* the rank that owns ghost node b of a distributed level (index into L->nbr), or -1 */
static int ghost_owner(bfmi_hier_level_t const* L, int32_t b) {
	for (int k = 0; k < L->n_nbr; k++) {
		if (b >= L->recv_begin[k] && b < L->recv_begin[k] + L->recv_count[k]) {
			return k;
		}
	}

	return -1;
}

bfmi_hier_t* bfmi_hier_build(bfmi_plan_t const* plan, double const* coords, bfmi_part_t const* part) {
	/* aggregate sizes: nodes per aggregate on the mesh level and on the levels above; the last level is solved by
	 * a dense inverse and may hold this many nodes (three unknowns each) */
	int64_t const ratio0 = env_i64("BFM_MG_RATIO0", 16);
	int64_t const ratio = env_i64("BFM_MG_RATIO", env_i64("BFM_MG_SMOOTH", BFMI_MG_SMOOTH_DEFAULT) != 0 ? 8 : 6);
	int64_t const dense_nodes = env_i64("BFM_MG_DENSE_NODES", 1024);
	int64_t const dense_limit = 2 * dense_nodes; /* pieces of bins can exceed the target */
	int64_t replicated_nodes = env_i64("BFM_MG_REPLICATED_NODES", 65536);
	bool const smooth = env_i64("BFM_MG_SMOOTH", BFMI_MG_SMOOTH_DEFAULT) != 0; /* smoothed aggregation (build_transfer) */

	int const world = part != NULL ? part->world : 1;
	int const rank = part != NULL ? part->rank : 0;

	replicated_nodes = replicated_nodes < dense_limit ? dense_limit : replicated_nodes; /* the dense level is always replicated */

	int64_t n_glob = part != NULL ? (int64_t) part->n_nodes : plan->nb; /* nodes of the current level over all ranks */

	if (ratio0 < 2 || ratio < 2 || n_glob < 4 * ratio0 || world > BFMG_DIST_MAX_RANKS) {
		return NULL;
	}

	bfmi_hier_t* const h = calloc(1, sizeof *h);

	if (h == NULL) {
		return NULL;
	}

	h->plan = (bfmi_plan_t*) plan; /* level 0 borrows its pattern */
	bfmi_plan_retain(h->plan);

	bfmi_hier_level_t* L = &h->level[0];

	L->n = plan->nb;
	L->dofs = 2;
	L->n_slices = plan->n_slices;
	L->n_slots = plan->n_slots;
	L->slice_off = plan->slice_off;
	L->row_len = plan->row_len;
	L->scol = plan->scol;
	L->diag_pos = plan->diag_pos;
	L->pos = (double*) coords;
	L->owns_pattern = false;
	L->row_lo = 0;
	L->row_hi = plan->nb;

	if (part != NULL) {
		L->distributed = true;
		L->row_lo = part->own_begin;
		L->row_hi = part->own_end;
		L->n_nbr = part->n_nbr;
		L->n_send = part->n_send;

		if (part->n_nbr > BFMG_DIST_MAX_RANKS) {
			goto fail;
		}

		memcpy(L->nbr, part->nbr, (size_t) part->n_nbr * sizeof(int32_t));
		memcpy(L->recv_begin, part->recv_begin, (size_t) part->n_nbr * sizeof(int32_t));
		memcpy(L->recv_count, part->recv_count, (size_t) part->n_nbr * sizeof(int32_t));
		memcpy(L->send_ptr, part->send_ptr, ((size_t) part->n_nbr + 1) * sizeof(int32_t));

		L->send_idx = malloc(((size_t) part->n_send + 1) * sizeof(int32_t));

		if (L->send_idx == NULL) {
			goto fail;
		}

		memcpy(L->send_idx, part->send_idx, (size_t) part->n_send * sizeof(int32_t));
	}

	h->n_levels = 1;

	for (int l = 0;; l++) {
		L = &h->level[l];

		if (l + 1 >= BFMG_MG_MAX_LEVELS) {
			goto fail;
		}

		int32_t const n_own = L->row_hi - L->row_lo;
		int64_t target = n_glob / (l == 0 ? ratio0 : ratio);

		/* close to the dense level: aim straight at it instead of leaving a tiny level in between */
		if (target <= 2 * dense_nodes) {
			target = target < dense_nodes ? target : dense_nodes;
		}

		if (target < 4) {
			goto fail; /* the same decision on every rank: n_glob is global */
		}

		if (L->distributed) { /* this rank's share of the aggregates */
			target = (int64_t) ((double) target * (double) n_own / (double) n_glob + 0.5);
			target = target < 1 ? 1 : target;
		}

		L->agg = malloc(((size_t) L->n + 1) * sizeof *L->agg);
		L->smoothed = smooth;

		if (L->agg == NULL) {
			goto fail;
		}

		double const t_agg = hier_clock();
		int32_t const n_agg = aggregate_level(L, target, l == 0 ? MIN_NODES_LEVEL0 : 1, L->agg);

		if (getenv("BFM_MG_VERBOSE") != NULL) {
			fprintf(stderr, "[hier] level %d: %d nodes (%d own) -> %d aggregates in %.1f ms\n", l, L->n, n_own, n_agg, hier_clock() - t_agg);
		}

		double const t_rest = hier_clock();
		bfmi_hier_level_t* const N = &h->level[l + 1];

		N->dofs = 3;
		N->owns_pattern = true;
		h->n_levels = l + 2;

		int64_t n_glob_next = n_agg;

		if (!L->distributed) {
			/* one GPU, or a replicated level: everything is here */

			if (n_agg < 4 || (int64_t) n_agg * 10 > n_glob * 7) {
				goto fail; /* no coarsening to speak of (or nothing to aggregate): the caller falls back */
			}

			L->n_coarse = n_agg;
			N->n = n_agg;
			N->row_lo = 0;
			N->row_hi = n_agg;
			N->pos = calloc((size_t) n_agg * 2 + 2, sizeof *N->pos);

			int32_t** const rows = calloc((size_t) n_agg + 1, sizeof *rows);
			int32_t* const lens = calloc((size_t) n_agg + 1, sizeof *lens);

			int rv = N->pos != NULL && rows != NULL && lens != NULL ? 0 : -1;

			double t_step[5] = {hier_clock(), 0, 0, 0, 0};

			rv = rv < 0 ? rv : build_transfer(L);
			t_step[1] = hier_clock();
			rv = rv < 0 ? rv : centroids(L, n_agg, 0, N->pos); /* after the transfer lists: summed aggregate by aggregate */
			t_step[2] = hier_clock();
			rv = rv < 0 ? rv : coarse_rows(L, 0, n_agg, rows, lens);
			t_step[3] = hier_clock();
			rv = rv < 0 ? rv : layout_level(N, n_agg, rows, lens);
			t_step[4] = hier_clock();

			if (getenv("BFM_MG_VERBOSE") != NULL && L->n > 100000) {
				fprintf(stderr, "[hier] level %d: transfer lists %.1f ms, reference points %.1f ms, symbolic P^T A P %.1f ms, SELL layout %.1f ms\n", l, t_step[1] - t_step[0], t_step[2] - t_step[1], t_step[3] - t_step[2], t_step[4] - t_step[3]);
			}

			for (int32_t I = 0; rows != NULL && I < n_agg; I++) {
				free(rows[I]);
			}

			free(rows);
			free(lens);

			if (rv < 0) {
				goto fail;
			}
		}

		else {
			/* several GPUs, distributed level: every rank has aggregated its own nodes */

			double const mine = n_agg;
			double counts[BFMG_DIST_MAX_RANKS];

			if (comm_allgather(&mine, 1, counts) < 0) {
				goto fail;
			}

			int64_t first = 0; /* global id of this rank's first aggregate */

			n_glob_next = 0;

			for (int r = 0; r < world; r++) {
				first += r < rank ? (int64_t) counts[r] : 0;
				n_glob_next += (int64_t) counts[r];
			}

			if (n_glob_next < 4 || n_glob_next * 10 > n_glob * 7 || n_glob_next > INT32_MAX / 4) {
				goto fail; /* the same decision on every rank */
			}

			bool const replicate = n_glob_next <= replicated_nodes;

			/* aggregate ids of the next level's numbering: global when it is replicated, owner-local otherwise */

			int32_t const shift = replicate ? (int32_t) first : 0;

			for (int32_t a = L->row_lo; a < L->row_hi && shift > 0; a++) {
				L->agg[a] += L->agg[a] >= 0 ? shift : 0;
			}

			double* const cen = calloc((size_t) n_agg * 2 + 2, sizeof *cen);
			double* const ids = calloc(((size_t) L->n + 1) * 2, sizeof *ids);
			double* const geo = calloc(((size_t) L->n + 1) * 2, sizeof *geo);

			int rv = cen != NULL && ids != NULL && geo != NULL ? 0 : -1;

			rv = rv < 0 ? rv : centroids(L, n_agg, shift, cen);

			for (int32_t a = L->row_lo; rv == 0 && a < L->row_hi; a++) {
				ids[2 * (size_t) a] = L->agg[a];
				geo[2 * (size_t) a + 0] = L->geom[2 * (size_t) a + 0];
				geo[2 * (size_t) a + 1] = L->geom[2 * (size_t) a + 1];
			}

			/* the aggregates and geometry of the ghost nodes come from their owners */

			rv = rv < 0 ? rv : comm_halo(L, ids);
			rv = rv < 0 ? rv : comm_halo(L, geo);

			if (rv < 0) {
				free(cen);
				free(ids);
				free(geo);
				goto fail;
			}

			for (int32_t b = 0; b < L->n; b++) {
				if (b < L->row_lo || b >= L->row_hi) {
					L->geom[2 * (size_t) b + 0] = (float) geo[2 * (size_t) b + 0];
					L->geom[2 * (size_t) b + 1] = (float) geo[2 * (size_t) b + 1];
				}
			}

			free(geo);

			if (replicate) {
				/* the next level is held by every rank in full, numbered globally (rank after rank) */

				for (int32_t b = 0; b < L->n; b++) {
					if (b < L->row_lo || b >= L->row_hi) {
						L->agg[b] = ghost_owner(L, b) >= 0 ? (int32_t) ids[2 * (size_t) b] : -1;
					}
				}

				free(ids);

				L->n_coarse = (int32_t) n_glob_next;
				L->gather_first = (int32_t) first;
				L->gather_count = n_agg;

				N->n = (int32_t) n_glob_next;
				N->row_lo = 0;
				N->row_hi = N->n;

				/* reference points and pattern rows of everybody's aggregates */

				double* all_pos = NULL;
				size_t per_rank[BFMG_DIST_MAX_RANKS];

				rv = comm_allgather_v(cen, (size_t) n_agg * 2, &all_pos, per_rank);
				free(cen);

				if (rv < 0 || build_transfer(L) < 0) {
					free(all_pos);
					goto fail;
				}

				N->pos = all_pos;

				int32_t** const rows = calloc((size_t) n_agg + 1, sizeof *rows);
				int32_t* const lens = calloc((size_t) n_agg + 1, sizeof *lens);

				rv = rows != NULL && lens != NULL ? coarse_rows(L, (int32_t) first, n_agg, rows, lens) : -1;

				size_t packed_len = 0;

				for (int32_t i = 0; rv == 0 && i < n_agg; i++) {
					packed_len += 1 + (size_t) lens[i];
				}

				double* const packed = malloc((packed_len + 1) * sizeof *packed);
				double* all_rows = NULL;

				if (rv == 0 && packed != NULL) {
					size_t at = 0;

					for (int32_t i = 0; i < n_agg; i++) {
						packed[at++] = lens[i];

						for (int32_t t = 0; t < lens[i]; t++) {
							packed[at++] = rows[i][t];
						}
					}

					rv = comm_allgather_v(packed, packed_len, &all_rows, per_rank);
				}

				else {
					rv = -1;
				}

				for (int32_t i = 0; rows != NULL && i < n_agg; i++) {
					free(rows[i]);
				}

				free(rows);
				free(lens);
				free(packed);

				if (rv < 0) {
					free(all_rows);
					goto fail;
				}

				/* unpack: rows arrive in global order (rank after rank, each rank's in its own order) */

				int32_t** const grows = calloc((size_t) N->n + 1, sizeof *grows);
				int32_t* const glens = calloc((size_t) N->n + 1, sizeof *glens);

				rv = grows != NULL && glens != NULL ? 0 : -1;

				size_t at = 0, total = 0;

				for (int r = 0; r < world; r++) {
					total += per_rank[r];
				}

				for (int32_t I = 0; rv == 0 && I < N->n; I++) {
					if (at >= total) {
						rv = -1;
						break;
					}

					glens[I] = (int32_t) all_rows[at++];
					grows[I] = malloc(((size_t) glens[I] + 1) * sizeof **grows);

					if (grows[I] == NULL || at + (size_t) glens[I] > total) {
						rv = -1;
						break;
					}

					for (int32_t t = 0; t < glens[I]; t++) {
						grows[I][t] = (int32_t) all_rows[at++];
					}
				}

				rv = rv < 0 ? rv : layout_level(N, N->n, grows, glens);

				for (int32_t I = 0; grows != NULL && I < N->n; I++) {
					free(grows[I]);
				}

				free(grows);
				free(glens);
				free(all_rows);

				if (rv < 0) {
					goto fail;
				}
			}

			else {
				/* the next level stays distributed: owned aggregates first, then the ghosts' aggregates grouped by
				 * owner (ascending rank), each group ascending by the owner's numbering - the order in which the
				 * owner will send them */

				int64_t* const keys = malloc(((size_t) L->n + 1) * sizeof *keys);
				size_t n_keys = 0;

				if (keys == NULL) {
					free(cen);
					free(ids);
					goto fail;
				}

				for (int32_t b = 0; b < L->n; b++) {
					if ((b < L->row_lo || b >= L->row_hi) && ids[2 * (size_t) b] >= 0) {
						int const k = ghost_owner(L, b);

						if (k >= 0) {
							keys[n_keys++] = (int64_t) k << 32 | (int64_t) ids[2 * (size_t) b];
						}
					}
				}

				qsort(keys, n_keys, sizeof *keys, cmp_i64);

				size_t uniq = 0;

				for (size_t i = 0; i < n_keys; i++) {
					if (i == 0 || keys[i] != keys[i - 1]) {
						keys[uniq++] = keys[i];
					}
				}

				n_keys = uniq;

				for (int32_t b = 0; b < L->n; b++) {
					if (b < L->row_lo || b >= L->row_hi) {
						int const k = ghost_owner(L, b);

						L->agg[b] = -1;

						if (k >= 0 && ids[2 * (size_t) b] >= 0) {
							int64_t const key = (int64_t) k << 32 | (int64_t) ids[2 * (size_t) b];
							int64_t const* const hit = bsearch(&key, keys, n_keys, sizeof *keys, cmp_i64);

							L->agg[b] = hit != NULL ? n_agg + (int32_t) (hit - keys) : -1;
						}
					}
				}

				free(ids);

				L->n_coarse = n_agg + (int32_t) n_keys;

				N->distributed = true;
				N->n = L->n_coarse;
				N->row_lo = 0;
				N->row_hi = n_agg;
				N->pos = calloc((size_t) N->n * 2 + 2, sizeof *N->pos);

				if (N->pos == NULL) {
					free(cen);
					free(keys);
					goto fail;
				}

				memcpy(N->pos, cen, (size_t) n_agg * 2 * sizeof *cen);
				free(cen);

				/* halo plan of the next level: what a neighbour ghosts of mine are the aggregates of the nodes it
				 * ghosts of mine; what I ghost of its are the aggregates of my ghosts of its */

				int32_t* const send = malloc(((size_t) L->n_send + 1) * sizeof *send);

				if (send == NULL) {
					free(keys);
					goto fail;
				}

				int32_t n_send = 0;

				N->n_nbr = 0;

				for (int k = 0; k < L->n_nbr; k++) {
					int32_t const begin = n_send;

					for (int32_t i = L->send_ptr[k]; i < L->send_ptr[k + 1]; i++) {
						int32_t const g = L->agg[L->send_idx[i]];

						if (g >= 0) {
							send[n_send++] = g;
						}
					}

					qsort(send + begin, (size_t) (n_send - begin), sizeof *send, cmp_i32);

					int32_t kept = begin;

					for (int32_t i = begin; i < n_send; i++) {
						if (i == begin || send[i] != send[i - 1]) {
							send[kept++] = send[i];
						}
					}

					n_send = kept;

					int32_t recv_first = -1, recv_count = 0;

					for (size_t i = 0; i < n_keys; i++) {
						if ((int) (keys[i] >> 32) == k) {
							recv_first = recv_first < 0 ? n_agg + (int32_t) i : recv_first;
							recv_count++;
						}
					}

					if (n_send == begin && recv_count == 0) {
						continue; /* no longer neighbours on the next level */
					}

					int const j = N->n_nbr++;

					N->nbr[j] = L->nbr[k];
					N->recv_begin[j] = recv_first < 0 ? n_agg : recv_first;
					N->recv_count[j] = recv_count;
					N->send_ptr[j] = begin;
					N->send_ptr[j + 1] = n_send;
				}

				N->n_send = n_send;
				N->send_idx = send;

				free(keys);

				int32_t** const rows = calloc((size_t) n_agg + 1, sizeof *rows);
				int32_t* const lens = calloc((size_t) n_agg + 1, sizeof *lens);

				rv = rows != NULL && lens != NULL ? 0 : -1;
				rv = rv < 0 ? rv : build_transfer(L);
				rv = rv < 0 ? rv : coarse_rows(L, 0, n_agg, rows, lens);
				rv = rv < 0 ? rv : layout_level(N, n_agg, rows, lens);

				for (int32_t I = 0; rows != NULL && I < n_agg; I++) {
					free(rows[I]);
				}

				free(rows);
				free(lens);

				if (rv < 0) {
					goto fail;
				}
			}
		}

		n_glob = n_glob_next;

		if (getenv("BFM_MG_VERBOSE") != NULL) {
			fprintf(stderr, "[hier] level %d: transfer + pattern of level %d (%s, %lld nodes over all ranks) in %.1f ms\n", l, l + 1, N->distributed ? "distributed" : "replicated", (long long) n_glob, hier_clock() - t_rest);
		}

		if (n_glob <= dense_limit) {
			if (N->distributed) {
				goto fail; /* cannot happen: replicated_nodes >= dense_limit */
			}

			break; /* N is the dense level */
		}
	}

	{
		bfmi_hier_level_t const* const last = &h->level[h->n_levels - 1];
		int32_t span = 0;

		for (int32_t I = 0; I < last->n; I++) {
			for (int32_t t = 0; t < last->row_len[I]; t++) {
				int32_t const d = abs(last->scol[slot_of(last->slice_off, I, t)] - I);
				span = d > span ? d : span;
			}
		}

		h->dense_span = span;
	}

	return h;

fail:

	bfmi_hier_free(h);
	return NULL;
}

/* ---- device mirrors ------------------------------------------------------------------------------------------ */

static int mirror(void** d_ptr, void const* src, size_t bytes, size_t* total) {
	if (bfmg_alloc(d_ptr, bytes > 0 ? bytes : 4) < 0) {
		return -1;
	}

	*total += bytes;
	return bytes > 0 ? bfmg_upload(*d_ptr, src, bytes) : 0;
}

int bfmi_hier_upload(bfmi_hier_t* h, bfmg_pattern_t const* pat0, size_t* h2d_bytes) {
	if (h->on_device) {
		return 0;
	}

	h->on_device = true; /* bfmi_hier_free releases whatever got allocated */

	bfmg_mg_t* const M = &h->dev;
	size_t bytes = 0;

	M->n_levels = h->n_levels;
	M->nc = (3 * h->level[h->n_levels - 1].n + 31) / 32 * 32;
	M->half_bw = 3 * h->dense_span + 2;

	for (int l = 0; l < h->n_levels; l++) {
		bfmi_hier_level_t* const L = &h->level[l];
		bfmg_mg_level_t* const d = &L->dev;

		d->n = L->n;
		d->dofs = L->dofs;
		d->n_slices = L->n_slices;
		d->n_slots = L->n_slots;
		d->row_lo = L->row_lo;
		d->row_hi = L->row_hi;
		d->distributed = L->distributed;
		d->gather_first = L->gather_first;
		d->gather_count = L->gather_count;
		d->n_nbr = L->n_nbr;
		d->n_send = L->n_send;

		for (int k = 0; k < L->n_nbr; k++) {
			d->nbr[k] = L->nbr[k];
			d->recv_begin[k] = L->recv_begin[k];
			d->recv_count[k] = L->recv_count[k];
			d->send_ptr[k] = L->send_ptr[k];
			d->send_ptr[k + 1] = L->send_ptr[k + 1];
		}

		if (L->distributed && mirror((void**) &d->send_idx, L->send_idx, (size_t) L->n_send * sizeof(int32_t), &bytes) < 0) {
			return -1;
		}

		if (l == 0) { /* the plan's pattern is on the device already */
			d->slice_off = pat0->slice_off;
			d->scol = pat0->scol;
			d->diag_pos = pat0->diag_pos;
			d->row_len = pat0->row_len;
		}

		else if (
			mirror((void**) &d->slice_off, L->slice_off, ((size_t) L->n_slices + 1) * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->scol, L->scol, (size_t) L->n_slots * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->diag_pos, L->diag_pos, (size_t) L->row_hi * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->row_len, L->row_len, (size_t) L->row_hi * sizeof(int32_t), &bytes) < 0
		) {
			return -1;
		}

		if (l + 1 < h->n_levels) {
			d->n_coarse = L->n_coarse;
			d->n_p = L->n_p;
			d->smoothed = L->smoothed ? 1 : 0;

			if (
				mirror((void**) &d->agg, L->agg, (size_t) L->n * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->geom, L->geom, (size_t) L->n * 2 * sizeof(float), &bytes) < 0 ||
				mirror((void**) &d->p_ptr, L->p_ptr, ((size_t) L->n + 1) * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->p_col, L->p_col, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_ptr, L->r_ptr, ((size_t) L->n_coarse + 1) * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_ent, L->r_ent, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_node, L->r_node, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0
			) {
				return -1;
			}
		}

		M->level[l] = *d;
	}

	if (h2d_bytes != NULL) {
		*h2d_bytes += bytes;
	}

	return 0;
}

/* ---- cache --------------------------------------------------------------------------------------------------
 * examples/benchmark.py-style loops call bfm_sim_run again and again on one mesh: the hierarchy and its device
 * mirrors are kept as long as mesh identity, connectivity and coordinates are the same. */

static bfmi_hier_t* cached;
static bool cached_none;
static bfmi_hier_t cached_key;

static uint64_t hash_coords(double const* coords, size_t count) {
	size_t const chunk = 1 << 16;
	size_t const n_chunks = (count + chunk - 1) / chunk;
	uint64_t total = 0x51ed270b7a2c3f11ull ^ count;
	uint64_t const* const bits = (uint64_t const*) coords; /* the bit patterns: -0.0 and 0.0 are different meshes here, harmlessly */

#pragma omp parallel for schedule(static) reduction(^ : total) if (n_chunks > 16)
	for (size_t c = 0; c < n_chunks; c++) {
		size_t const end = (c + 1) * chunk < count ? (c + 1) * chunk : count;
		uint64_t h0 = 0xcbf29ce484222325ull + c, h1 = 0x84222325cbf29ce4ull ^ c, h2 = 0x9e3779b97f4a7c15ull + 3 * c, h3 = 0xc2b2ae3d27d4eb4full ^ (c << 7);
		size_t i = c * chunk;

		for (; i + 4 <= end; i += 4) { /* four independent chains: memory speed, not multiplier latency */
			h0 = (h0 ^ bits[i + 0]) * 0x100000001b3ull;
			h1 = (h1 ^ bits[i + 1]) * 0x9e3779b97f4a7c15ull;
			h2 = (h2 ^ bits[i + 2]) * 0xc2b2ae3d27d4eb4full;
			h3 = (h3 ^ bits[i + 3]) * 0x165667b19e3779f9ull;
			h0 ^= h0 >> 31, h1 ^= h1 >> 29, h2 ^= h2 >> 27, h3 ^= h3 >> 33;
		}

		for (; i < end; i++) {
			h0 = (h0 ^ bits[i]) * 0x100000001b3ull;
			h0 ^= h0 >> 31;
		}

		uint64_t hsh = (h0 * 31 + h1) * 0x100000001b3ull;

		hsh = ((hsh ^ (hsh >> 32)) * 31 + h2) * 0x9e3779b97f4a7c15ull;
		hsh = ((hsh ^ (hsh >> 29)) * 31 + h3) * 0xc2b2ae3d27d4eb4full;

		total ^= (hsh ^ (hsh >> 31)) * (2 * c + 1);
	}

	return total;
}

static uint64_t settings_hash(void) {
	return (uint64_t) env_i64("BFM_MG_RATIO0", 16) * 1000003u + (uint64_t) env_i64("BFM_MG_RATIO", env_i64("BFM_MG_SMOOTH", BFMI_MG_SMOOTH_DEFAULT) != 0 ? 8 : 6) * 10007u + (uint64_t) env_i64("BFM_MG_DENSE_NODES", 1024) + (uint64_t) env_i64("BFM_MG_REPLICATED_NODES", 65536) * 7919u + (uint64_t) env_i64("BFM_MG_SMOOTH", BFMI_MG_SMOOTH_DEFAULT) * 104729u;
}

static bool same_key(bfmi_hier_t const* h, bfmi_plan_t const* plan, uint64_t coords_hash, int rank, int world) {
	return h->key_plan == plan && h->key_nodes == (size_t) plan->nb && h->elems_hash == plan->elems_hash && h->coords_hash == coords_hash && h->rank == rank && h->world == world && h->settings == settings_hash();
}

static void set_key(bfmi_hier_t* h, bfmi_plan_t const* plan, uint64_t coords_hash, int rank, int world) {
	h->key_plan = plan;
	h->key_nodes = (size_t) plan->nb;
	h->elems_hash = plan->elems_hash;
	h->coords_hash = coords_hash;
	h->rank = rank;
	h->world = world;
	h->settings = settings_hash();
}

void bfmi_hier_release(bfmi_hier_t* h) {
	if (h != NULL && __atomic_sub_fetch(&h->refs, 1, __ATOMIC_ACQ_REL) == 0) {
		bfmi_hier_free(h);
	}
}

bfmi_hier_t* bfmi_hier_for_plan(bfmi_plan_t* plan, double const* coords, bfmi_part_t const* part) {
	int const rank = part != NULL ? part->rank : 0;
	int const world = part != NULL ? part->world : 1;
	uint64_t const coords_hash = hash_coords(coords, (size_t) plan->nb * 2);

	if (cached != NULL && same_key(cached, plan, coords_hash, rank, world)) {
		__atomic_add_fetch(&cached->refs, 1, __ATOMIC_RELAXED);
		return cached;
	}

	if (cached_none && same_key(&cached_key, plan, coords_hash, rank, world)) {
		return NULL;
	}

	bfmi_hier_t* const h = bfmi_hier_build(plan, coords, part);

	if (h == NULL) {
		cached_none = true;
		set_key(&cached_key, plan, coords_hash, rank, world);
		return NULL;
	}

	h->refs = 1; /* the cache's */
	set_key(h, plan, coords_hash, rank, world);

	bfmi_hier_release(cached);
	cached = h;

	__atomic_add_fetch(&h->refs, 1, __ATOMIC_RELAXED); /* the caller's */
	return h;
}

void bfmi_hier_forget(bfmi_plan_t const* plan) {
	if (cached != NULL && cached->key_plan == plan) {
		bfmi_hier_release(cached);
		cached = NULL;
	}

	if (cached_none && cached_key.key_plan == plan) {
		cached_none = false;
	}
}

/* ---- introspection for tests (bfm_b200.h): host-only, needs no device ------------------------------------------- */

static bfmi_hier_t* hier_of(bfm_mesh_t* mesh, bfmi_plan_t** plan_out) {
	bfmi_plan_t* const plan = bfmi_plan_for_mesh(mesh->state, mesh);

	if (plan == NULL) {
		return NULL;
	}

	bfmi_hier_t* const h = bfmi_hier_for_plan(plan, mesh->coords, NULL);

	*plan_out = plan;
	return h;
}

int bfmx_hier_info(bfm_mesh_t* mesh, bfmx_hier_info_t* info) {
	bfmi_plan_t* plan = NULL;
	bfmi_hier_t* const h = hier_of(mesh, &plan);

	memset(info, 0, sizeof *info);

	if (h != NULL) {
		info->n_levels = h->n_levels;

		for (int l = 0; l < h->n_levels; l++) {
			info->n_nodes[l] = h->level[l].n;
			info->n_slots[l] = h->level[l].n_slots;
			info->n_entries[l] = l + 1 < h->n_levels ? h->level[l].n_p : 0;
		}

		info->smoothed = h->level[0].smoothed ? 1 : 0;
	}

	bfmi_hier_release(h);
	bfmi_plan_release(plan);

	return plan != NULL ? 0 : -1;
}

int bfmx_hier_level(bfm_mesh_t* mesh, int level, int32_t* aggregate, float* geometry, int32_t* pattern_rowptr, int32_t* pattern_col) {
	bfmi_plan_t* plan = NULL;
	bfmi_hier_t* const h = hier_of(mesh, &plan);
	int rv = -1;

	if (h != NULL && level >= 0 && level < h->n_levels) {
		bfmi_hier_level_t const* const L = &h->level[level];

		rv = 0;

		if (level + 1 < h->n_levels) {
			if (aggregate != NULL) {
				memcpy(aggregate, L->agg, (size_t) L->n * sizeof *aggregate);
			}

			if (geometry != NULL) {
				memcpy(geometry, L->geom, (size_t) L->n * 2 * sizeof *geometry);
			}

		}

		else if (aggregate != NULL || geometry != NULL) {
			rv = -1; /* the last level has nothing above it */
		}

		if (pattern_rowptr != NULL) {
			int32_t at = 0;

			for (int32_t a = 0; a < L->n; a++) {
				pattern_rowptr[a] = at;

				for (int32_t t = 0; t < L->row_len[a]; t++, at++) {
					if (pattern_col != NULL) {
						pattern_col[at] = L->scol[slot_of(L->slice_off, a, t)];
					}
				}
			}

			pattern_rowptr[L->n] = at;
		}
	}

	bfmi_hier_release(h);
	bfmi_plan_release(plan);

	return rv;
}
