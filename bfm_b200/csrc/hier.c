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
 * values are computed on the device for every solve because they carry the matrix's diagonal scaling.  The
 * next level's operator P^T A P is computed on the device by probing: coarse nodes that share no row of
 * P^T A P get the same colour, so one product with the sum of one mode over one colour yields one column
 * per coarse node of that colour.  Pattern of P^T A P and the colouring are computed here, symbolically.
 *
 * Everything is deterministic: ids follow the smallest member, sets are sorted, colours are greedy in id
 * order.
 */
#include "internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SLICE 32
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

	if (L->owns_owner) {
		free((void*) L->owner);
	}

	free(L->agg);
	free(L->geom);
	free(L->p_ptr);
	free(L->p_col);
	free(L->r_ptr);
	free(L->r_ent);
	free(L->r_node);
	free(L->color);

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
	bfmg_free(d->color);
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

static int32_t aggregate_level(bfmi_hier_level_t const* L, int64_t target, int32_t min_nodes, int32_t* agg) {
	int32_t const n = L->n;
	double const* const pos = L->pos;

	double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;

	for (int32_t a = 0; a < n; a++) {
		double const x = pos[2 * (size_t) a + 0];
		double const y = pos[2 * (size_t) a + 1];

		if (!(x == x) || !(y == y) || isinf(x) || isinf(y)) {
			return 0;
		}

		x0 = x < x0 ? x : x0, x1 = x > x1 ? x : x1;
		y0 = y < y0 ? y : y0, y1 = y > y1 ? y : y1;
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
		int64_t bx = (int64_t) ((pos[2 * (size_t) a + 0] - x0) / wx * (double) nbx);
		int64_t by = (int64_t) ((pos[2 * (size_t) a + 1] - y0) / wy * (double) nby);

		bx = bx >= nbx ? nbx - 1 : (bx < 0 ? 0 : bx);
		by = by >= nby ? nby - 1 : (by < 0 ? 0 : by);

		bin[a] = (int32_t) (by * nbx + bx);
		parent[a] = a;
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
	 * truss-like geometry a bin can cut through members that do not touch and each becomes its own aggregate */

	for (int32_t a = 0; a < n; a++) {
		for (int32_t t = 0; t < L->row_len[a]; t++) {
			int32_t const b = L->scol[slot_of(L->slice_off, a, t)];

			if (b < a && bin[b] == bin[a] && (L->owner == NULL || L->owner[a] == L->owner[b])) {
				int32_t ra, rb;

				FIND(a, ra);
				FIND(b, rb);

				if (ra != rb) {
					if (ra < rb) {
						parent[rb] = ra;
					}

					else {
						parent[ra] = rb;
					}
				}
			}
		}
	}

	/* pieces below min_nodes join a piece they are coupled to (a few passes: a chain of tiny pieces needs more
	 * than one) */

	for (int pass = 0; pass < 4 && min_nodes > 1; pass++) {
		memset(size, 0, ((size_t) n + 1) * sizeof *size);

		for (int32_t a = 0; a < n; a++) {
			int32_t r;
			FIND(a, r);
			size[r]++;
		}

		bool changed = false;

		for (int32_t a = 0; a < n; a++) {
			int32_t ra;
			FIND(a, ra);

			if (size[ra] >= min_nodes) {
				continue;
			}

			for (int32_t t = 0; t < L->row_len[a]; t++) {
				int32_t const b = L->scol[slot_of(L->slice_off, a, t)];
				int32_t rb;

				if (b == a || (L->owner != NULL && L->owner[a] != L->owner[b])) {
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

	memset(size, 0, ((size_t) n + 1) * sizeof *size);

	for (int32_t a = 0; a < n; a++) {
		int32_t r;
		FIND(a, r);
		size[r]++;
	}

	/* compact: ids in order of each piece's smallest node (its root); pieces still too small are left out */

	int32_t n_agg = 0;

	for (int32_t a = 0; a < n; a++) {
		if (parent[a] == a) {
			bin[a] = size[a] >= min_nodes ? n_agg++ : -1; /* bin[] is free now: reuse as root -> id */
		}
	}

	for (int32_t a = 0; a < n; a++) {
		int32_t r;
		FIND(a, r);
		agg[a] = bin[r];
	}

#undef FIND

	free(bin);
	free(parent);
	free(size);

	return n_agg;
}

/* ---- the structures between level l and level l + 1 -------------------------------------------------------- */

/* CSR of the prolongator (one entry per node: its aggregate) and its transpose */
static int build_transfer(bfmi_hier_level_t* L) {
	int32_t const n = L->n;
	int32_t const nc = L->n_coarse;

	L->p_ptr = malloc(((size_t) n + 1) * sizeof *L->p_ptr);
	L->r_ptr = calloc((size_t) nc + 2, sizeof *L->r_ptr);

	if (L->p_ptr == NULL || L->r_ptr == NULL) {
		return -1;
	}

	int32_t n_p = 0;

	for (int32_t a = 0; a < n; a++) {
		L->p_ptr[a] = n_p;
		n_p += L->agg[a] >= 0;
	}

	L->p_ptr[n] = n_p;
	L->n_p = n_p;

	L->p_col = malloc(((size_t) n_p + 1) * sizeof *L->p_col);
	L->r_ent = malloc(((size_t) n_p + 1) * sizeof *L->r_ent);
	L->r_node = malloc(((size_t) n_p + 1) * sizeof *L->r_node);

	if (L->p_col == NULL || L->r_ent == NULL || L->r_node == NULL) {
		return -1;
	}

	for (int32_t a = 0; a < n; a++) {
		if (L->agg[a] >= 0) {
			L->p_col[L->p_ptr[a]] = L->agg[a];
			L->r_ptr[L->agg[a] + 2]++;
		}
	}

	for (int32_t g = 0; g < nc; g++) {
		L->r_ptr[g + 2] += L->r_ptr[g + 1];
	}

	for (int32_t a = 0; a < n; a++) { /* ascending nodes inside every coarse node's list */
		for (int32_t e = L->p_ptr[a]; e < L->p_ptr[a + 1]; e++) {
			int32_t const at = L->r_ptr[L->p_col[e] + 1]++;

			L->r_ent[at] = e;
			L->r_node[at] = a;
		}
	}

	return 0;
}

static int cmp_i32(void const* a, void const* b) {
	int32_t const x = *(int32_t const*) a;
	int32_t const y = *(int32_t const*) b;
	return x < y ? -1 : x > y;
}

/* pattern of P^T A P as a SELL-32 node pattern of the next level: coarse node I couples to J when some node in
 * the support of column I of P is coupled (through A) to a node in the support of column J */
static int build_coarse_pattern(bfmi_hier_level_t const* L, bfmi_hier_level_t* N) {
	int32_t const nc = L->n_coarse;

	N->n = nc;
	N->owns_pattern = true;
	N->row_len = calloc((size_t) nc + 1, sizeof *N->row_len);

	int64_t* const row_ptr = calloc((size_t) nc + 2, sizeof *row_ptr);

	if (N->row_len == NULL || row_ptr == NULL) {
		free(row_ptr);
		return -1;
	}

	int const n_threads =
#ifdef _OPENMP
		nc > 20000 ? omp_get_max_threads() : 1;
#else
		1;
#endif

	int32_t* const marks = malloc((size_t) n_threads * ((size_t) nc + 1) * sizeof *marks);
	int32_t** const rows = calloc((size_t) nc + 1, sizeof *rows); /* per coarse node: its sorted column list */

	if (marks == NULL || rows == NULL) {
		free(marks);
		free(rows);
		free(row_ptr);
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
		for (int32_t I = 0; I < nc; I++) {
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

			qsort(buf, (size_t) cnt, sizeof *buf, cmp_i32);

			rows[I] = malloc((size_t) cnt * sizeof **rows);

			if (rows[I] == NULL) {
				failed = true;
				continue;
			}

			memcpy(rows[I], buf, (size_t) cnt * sizeof *buf);
			N->row_len[I] = cnt;
		}

		free(buf);
	}

	free(marks);

	int rv = -1;

	if (failed) {
		goto done;
	}

	N->n_slices = (nc + SLICE - 1) / SLICE;
	N->slice_off = malloc(((size_t) N->n_slices + 1) * sizeof *N->slice_off);

	if (N->slice_off == NULL) {
		goto done;
	}

	{
		int64_t off = 0;

		for (int32_t s = 0; s < N->n_slices; s++) {
			int32_t longest = 0;

			for (int32_t a = s * SLICE; a < nc && a < (s + 1) * SLICE; a++) {
				longest = N->row_len[a] > longest ? N->row_len[a] : longest;
			}

			N->slice_off[s] = (int32_t) off;
			off += (int64_t) longest * SLICE;

			if (off > INT32_MAX / 9) { /* nine value planes are indexed with 32-bit slots */
				goto done;
			}
		}

		N->slice_off[N->n_slices] = (int32_t) off;
		N->n_slots = off;
	}

	N->scol = malloc(((size_t) N->n_slots + 1) * sizeof *N->scol);
	N->diag_pos = malloc(((size_t) nc + 1) * sizeof *N->diag_pos);

	if (N->scol == NULL || N->diag_pos == NULL) {
		goto done;
	}

	/* padding slots point at their own row (clamped) and carry zeros */

	for (int32_t s = 0; s < N->n_slices; s++) {
		for (int32_t slot = N->slice_off[s]; slot < N->slice_off[s + 1]; slot++) {
			int32_t const a = s * SLICE + (slot - N->slice_off[s]) % SLICE;
			N->scol[slot] = a < nc ? a : nc - 1;
		}
	}

	for (int32_t I = 0; I < nc; I++) {
		for (int32_t t = 0; t < N->row_len[I]; t++) {
			int64_t const slot = slot_of(N->slice_off, I, t);

			N->scol[slot] = rows[I][t];

			if (rows[I][t] == I) {
				N->diag_pos[I] = (int32_t) slot;
			}
		}
	}

	rv = 0;

done:

	for (int32_t I = 0; I < nc; I++) {
		free(rows[I]);
	}

	free(rows);
	free(row_ptr);

	return rv;
}

/* greedy distance-2 colouring of the next level's pattern: I differs from every J that shares a row with it
 * (J in row K and I in row K for some K; the pattern is symmetric), so that the columns probed together never
 * meet in a row */
static int color_level(bfmi_hier_level_t* L, bfmi_hier_level_t const* N) {
	int32_t const nc = N->n;

	L->color = malloc(((size_t) nc + 1) * sizeof *L->color);

	int32_t* const mark = malloc(((size_t) nc + 1) * sizeof *mark); /* indexed by colour: never more colours than nodes */

	if (L->color == NULL || mark == NULL) {
		free(mark);
		return -1;
	}

	for (int32_t I = 0; I < nc; I++) {
		L->color[I] = -1;
		mark[I] = -1;
	}

	int32_t n_colors = 0;

	for (int32_t I = 0; I < nc; I++) {
		for (int32_t t = 0; t < N->row_len[I]; t++) {
			int32_t const K = N->scol[slot_of(N->slice_off, I, t)];

			for (int32_t u = 0; u < N->row_len[K]; u++) {
				int32_t const J = N->scol[slot_of(N->slice_off, K, u)];

				if (J != I && L->color[J] >= 0) {
					mark[L->color[J]] = I;
				}
			}
		}

		int32_t col = 0;

		while (mark[col] == I) {
			col++;
		}

		L->color[I] = col;
		n_colors = col + 1 > n_colors ? col + 1 : n_colors;
	}

	L->n_colors = n_colors;

	free(mark);
	return 0;
}

/* ---- the hierarchy ---------------------------------------------------------------------------------------- */

static int64_t env_i64(char const* name, int64_t fallback) {
	char const* const env = getenv(name);
	return env != NULL && env[0] != 0 ? atoll(env) : fallback;
}

bfmi_hier_t* bfmi_hier_build(bfmi_plan_t const* plan, double const* coords, int32_t const* owner) {
	/* aggregate sizes: nodes per aggregate on the mesh level and on the levels above; the last level is solved by
	 * a dense inverse and may hold this many nodes (three unknowns each) */
	int64_t const ratio0 = env_i64("BFM_MG_RATIO0", 12);
	int64_t const ratio = env_i64("BFM_MG_RATIO", 6);
	int64_t const dense_nodes = env_i64("BFM_MG_DENSE_NODES", 640);
	int64_t const dense_limit = 2 * dense_nodes; /* pieces of bins can exceed the target */

	if (ratio0 < 2 || ratio < 2 || plan->nb < 4 * ratio0) {
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
	L->owner = owner;
	L->owns_pattern = false;

	h->n_levels = 1;

	for (int l = 0;; l++) {
		L = &h->level[l];

		if (l + 1 >= BFMG_MG_MAX_LEVELS) {
			goto fail;
		}

		int64_t target = L->n / (l == 0 ? ratio0 : ratio);

		/* close to the dense level: aim straight at it instead of leaving a tiny level in between */
		if (target <= 2 * dense_nodes) {
			target = target < dense_nodes ? target : dense_nodes;
		}

		if (target < 4) {
			goto fail;
		}

		L->agg = malloc(((size_t) L->n + 1) * sizeof *L->agg);

		if (L->agg == NULL) {
			goto fail;
		}

		int32_t const n_agg = aggregate_level(L, target, l == 0 ? MIN_NODES_LEVEL0 : 1, L->agg);

		if (n_agg < 4 || (int64_t) n_agg * 10 > (int64_t) L->n * 7) {
			goto fail; /* no coarsening to speak of (or nothing to aggregate): the caller falls back */
		}

		L->n_coarse = n_agg;

		bfmi_hier_level_t* const N = &h->level[l + 1];

		N->dofs = 3;
		h->n_levels = l + 2;

		/* reference points of the next level: centroids of the aggregates; geometry of P relative to them */

		N->pos = calloc((size_t) n_agg * 2 + 2, sizeof *N->pos);
		L->geom = malloc(((size_t) L->n + 1) * 2 * sizeof *L->geom);

		int32_t* const count = calloc((size_t) n_agg + 1, sizeof *count);

		if (N->pos == NULL || L->geom == NULL || count == NULL) {
			free(count);
			N->owns_pattern = true; /* so that level_free releases N->pos */
			goto fail;
		}

		N->owns_pattern = true;

		for (int32_t a = 0; a < L->n; a++) { /* fixed order: identical wherever it is computed */
			int32_t const g = L->agg[a];

			if (g >= 0) {
				N->pos[2 * (size_t) g + 0] += L->pos[2 * (size_t) a + 0];
				N->pos[2 * (size_t) g + 1] += L->pos[2 * (size_t) a + 1];
				count[g]++;
			}
		}

		for (int32_t g = 0; g < n_agg; g++) {
			N->pos[2 * (size_t) g + 0] /= (double) count[g];
			N->pos[2 * (size_t) g + 1] /= (double) count[g];
		}

		free(count);

		for (int32_t a = 0; a < L->n; a++) {
			int32_t const g = L->agg[a];

			L->geom[2 * (size_t) a + 0] = g >= 0 ? (float) (L->pos[2 * (size_t) a + 0] - N->pos[2 * (size_t) g + 0]) : 0;
			L->geom[2 * (size_t) a + 1] = g >= 0 ? (float) (L->pos[2 * (size_t) a + 1] - N->pos[2 * (size_t) g + 1]) : 0;
		}

		if (owner != NULL) {
			/* an aggregate belongs to the rank of its members (they all share one) */
			int32_t* const up = malloc(((size_t) n_agg + 1) * sizeof *up);

			if (up == NULL) {
				goto fail;
			}

			for (int32_t a = L->n - 1; a >= 0; a--) {
				if (L->agg[a] >= 0) {
					up[L->agg[a]] = L->owner[a];
				}
			}

			N->owner = up;
			N->owns_owner = true;
		}

		if (build_transfer(L) < 0 || build_coarse_pattern(L, N) < 0 || color_level(L, N) < 0) {
			goto fail;
		}

		if (n_agg <= dense_limit) {
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

		if (l == 0) { /* the plan's pattern is on the device already */
			d->slice_off = pat0->slice_off;
			d->scol = pat0->scol;
			d->diag_pos = pat0->diag_pos;
			d->row_len = pat0->row_len;
		}

		else if (
			mirror((void**) &d->slice_off, L->slice_off, ((size_t) L->n_slices + 1) * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->scol, L->scol, (size_t) L->n_slots * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->diag_pos, L->diag_pos, (size_t) L->n * sizeof(int32_t), &bytes) < 0 ||
			mirror((void**) &d->row_len, L->row_len, (size_t) L->n * sizeof(int32_t), &bytes) < 0
		) {
			return -1;
		}

		if (l + 1 < h->n_levels) {
			d->n_coarse = L->n_coarse;
			d->n_p = L->n_p;
			d->n_colors = L->n_colors;

			if (
				mirror((void**) &d->agg, L->agg, (size_t) L->n * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->geom, L->geom, (size_t) L->n * 2 * sizeof(float), &bytes) < 0 ||
				mirror((void**) &d->p_ptr, L->p_ptr, ((size_t) L->n + 1) * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->p_col, L->p_col, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_ptr, L->r_ptr, ((size_t) L->n_coarse + 1) * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_ent, L->r_ent, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->r_node, L->r_node, (size_t) L->n_p * sizeof(int32_t), &bytes) < 0 ||
				mirror((void**) &d->color, L->color, (size_t) L->n_coarse * sizeof(int32_t), &bytes) < 0
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

#pragma omp parallel for schedule(static) reduction(^ : total) if (n_chunks > 16)
	for (size_t c = 0; c < n_chunks; c++) {
		size_t const end = (c + 1) * chunk < count ? (c + 1) * chunk : count;
		uint64_t hsh = 0xcbf29ce484222325ull + c;

		for (size_t i = c * chunk; i < end; i++) {
			uint64_t bits;
			memcpy(&bits, &coords[i], sizeof bits);
			hsh = (hsh ^ bits) * 0x100000001b3ull;
			hsh ^= hsh >> 31;
		}

		total ^= hsh * (2 * c + 1);
	}

	return total;
}

static uint64_t settings_hash(void) {
	return (uint64_t) env_i64("BFM_MG_RATIO0", 12) * 1000003u + (uint64_t) env_i64("BFM_MG_RATIO", 6) * 10007u + (uint64_t) env_i64("BFM_MG_DENSE_NODES", 640);
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

bfmi_hier_t* bfmi_hier_for_plan(bfmi_plan_t* plan, double const* coords, int32_t const* owner, int rank, int world) {
	uint64_t const coords_hash = hash_coords(coords, (size_t) plan->nb * 2);

	if (cached != NULL && same_key(cached, plan, coords_hash, rank, world)) {
		__atomic_add_fetch(&cached->refs, 1, __ATOMIC_RELAXED);
		return cached;
	}

	if (cached_none && same_key(&cached_key, plan, coords_hash, rank, world)) {
		return NULL;
	}

	bfmi_hier_t* const h = bfmi_hier_build(plan, coords, owner);

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

	bfmi_hier_t* const h = bfmi_hier_for_plan(plan, mesh->coords, NULL, 0, 1);

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
			info->n_colors[l] = l + 1 < h->n_levels ? h->level[l].n_colors : 0;
			info->n_slots[l] = h->level[l].n_slots;
		}
	}

	bfmi_hier_release(h);
	bfmi_plan_release(plan);

	return plan != NULL ? 0 : -1;
}

int bfmx_hier_level(bfm_mesh_t* mesh, int level, int32_t* aggregate, float* geometry, int32_t* color, int32_t* pattern_rowptr, int32_t* pattern_col) {
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

			if (color != NULL) {
				memcpy(color, L->color, (size_t) L->n_coarse * sizeof *color);
			}
		}

		else if (aggregate != NULL || geometry != NULL || color != NULL) {
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
