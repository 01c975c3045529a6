/*
 * coarse.c - host side of the solver's coarse level: node aggregates and their colouring.
 *
 * The reference solves with a direct band LU (matrix.c:253-404); this build solves with conjugate
 * gradients, whose iteration count with a diagonal preconditioner grows like 1/h (46 000 iterations on
 * the 8 M-DOF plate, measured).  The coarse level turns that into ~30 * (H/h): the preconditioner is
 *
 *     M^-1 = I + W E^-1 W^T        (in the Jacobi-scaled variables of solver.cu),   E = W^T A^ W
 *
 * where W holds, per aggregate of nodes, the three rigid-body modes of plane elasticity (two
 * translations, one rotation about the aggregate's centroid) - the near-null space that Jacobi cannot
 * see.  It changes how fast CG converges, not what it converges to: the stopping test stays
 * ||b - A x|| <= 1e-12 ||b||.
 *
 * Aggregates are the non-empty cells of a uniform grid of bins over the mesh's bounding box (bins with
 * fewer than three nodes are merged into a neighbour, so that the three modes of every aggregate are
 * independent).  They are defined on the GLOBAL mesh from coordinates alone, so every rank of a
 * multi-GPU job computes the same aggregates without communication.
 *
 * E is computed on the device by probing (coarse.cu): aggregates that are adjacent to (or equal to) a
 * common aggregate get different colours (distance-2 colouring of the aggregate graph, greedy), so one
 * SpMV with the sum of one mode over one colour yields, after restriction, one column of E per
 * aggregate of that colour.
 */
#include "internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MIN_NODES_PER_AGGREGATE 3

static int cmp_u64(void const* a, void const* b) {
	uint64_t const x = *(uint64_t const*) a;
	uint64_t const y = *(uint64_t const*) b;
	return x < y ? -1 : x > y;
}

void bfmi_coarse_free(bfmi_coarse_t* c) {
	if (c == NULL) {
		return;
	}

	bfmg_free(c->dev.agg);
	bfmg_free(c->dev.wgeom);
	bfmg_free(c->dev.agg_ptr);
	bfmg_free(c->dev.agg_nodes);
	bfmg_free(c->dev.color_nbr);
	bfmg_free(c->dev.color);

	free(c->agg);
	free(c->wgeom);
	free(c->agg_ptr);
	free(c->agg_nodes);
	free(c->color);
	free(c->color_nbr);
	free(c);
}

/* aggregate of every GLOBAL node, compacted to 0 .. n_agg - 1; returns n_agg (0: no coarse level) */
static int32_t aggregate_nodes(bfm_mesh_t const* mesh, int32_t target, int32_t* node_agg) {
	size_t const nn = mesh->n_nodes;

	double x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;

	for (size_t a = 0; a < nn; a++) {
		double const x = mesh->coords[2 * a + 0];
		double const y = mesh->coords[2 * a + 1];

		if (!(x == x) || !(y == y) || isinf(x) || isinf(y)) {
			return 0;
		}

		x0 = x < x0 ? x : x0, x1 = x > x1 ? x : x1;
		y0 = y < y0 ? y : y0, y1 = y > y1 ? y : y1;
	}

	double const wx = x1 - x0;
	double const wy = y1 - y0;

	if (!(wx > 0) || !(wy > 0)) {
		return 0; /* degenerate geometry: nothing to bin */
	}

	double const side = sqrt(wx * wy / (double) target);
	int64_t nbx = (int64_t) floor(wx / side + 0.5);
	int64_t nby = (int64_t) floor(wy / side + 0.5);

	nbx = nbx < 1 ? 1 : nbx;
	nby = nby < 1 ? 1 : nby;

	while (nbx * nby > 4 * (int64_t) target) { /* extreme aspect ratios */
		nbx > nby ? nbx-- : nby--;
	}

	int64_t const n_bins = nbx * nby;
	int32_t* const count = calloc((size_t) n_bins, sizeof *count);
	int32_t* const remap = malloc((size_t) n_bins * sizeof *remap);

	if (count == NULL || remap == NULL) {
		free(count);
		free(remap);
		return 0;
	}

	for (size_t a = 0; a < nn; a++) {
		int64_t bx = (int64_t) ((mesh->coords[2 * a + 0] - x0) / wx * (double) nbx);
		int64_t by = (int64_t) ((mesh->coords[2 * a + 1] - y0) / wy * (double) nby);

		bx = bx >= nbx ? nbx - 1 : (bx < 0 ? 0 : bx);
		by = by >= nby ? nby - 1 : (by < 0 ? 0 : by);

		node_agg[a] = (int32_t) (by * nbx + bx);
		count[node_agg[a]]++;
	}

	/* Aggregates = the connected pieces of each bin: nodes of one bin joined through elements that lie inside
	 * it.  On a solid mesh a bin is one piece; on truss-like geometry (bridge-dam.obj) a bin can cut through
	 * several members that do not touch, and one set of rigid-body modes for all of them is a poor coarse
	 * function (measured: 535 iterations with plain bins).  Union-find, smaller root wins, so the
	 * result does not depend on traversal order. */

	size_t const kind = mesh->kind;
	int32_t* const parent = malloc((nn + 1) * sizeof *parent);
	int32_t* const size = calloc(nn + 1, sizeof *size);

	if (parent == NULL || size == NULL) {
		free(parent);
		free(size);
		free(count);
		free(remap);
		return 0;
	}

	for (size_t a = 0; a < nn; a++) {
		parent[a] = (int32_t) a;
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

#define UNION(u, v)                                \
	do {                                           \
		int32_t ru_, rv_;                          \
		FIND((u), ru_);                            \
		FIND((v), rv_);                            \
		if (ru_ != rv_) {                          \
			if (ru_ < rv_) parent[rv_] = ru_;      \
			else parent[ru_] = rv_;                \
		}                                          \
	} while (0)

	for (size_t e = 0; e < mesh->n_elems; e++) {
		size_t const* const el = &mesh->elems[e * kind];

		for (size_t j = 1; j < kind; j++) {
			for (size_t k = 0; k < j; k++) {
				if (el[j] < nn && el[k] < nn && node_agg[el[j]] == node_agg[el[k]]) {
					UNION((int32_t) el[j], (int32_t) el[k]);
				}
			}
		}
	}

	/* pieces too small to carry three independent modes join a piece they touch through an element (a few
	 * passes: a chain of tiny pieces needs more than one); what is still small afterwards - isolated nodes,
	 * islands of one or two nodes - joins the largest piece of its bin, or of the nearest bin that has one */

	for (int pass = 0; pass < 4; pass++) {
		memset(size, 0, (nn + 1) * sizeof *size);

		for (size_t a = 0; a < nn; a++) {
			int32_t r;
			FIND((int32_t) a, r);
			size[r]++;
		}

		bool changed = false;

		for (size_t e = 0; e < mesh->n_elems; e++) {
			size_t const* const el = &mesh->elems[e * kind];

			for (size_t j = 1; j < kind; j++) {
				int32_t r0, rj;

				if (el[0] >= nn || el[j] >= nn) {
					continue;
				}

				FIND((int32_t) el[0], r0);
				FIND((int32_t) el[j], rj);

				if (r0 != rj && (size[r0] < MIN_NODES_PER_AGGREGATE || size[rj] < MIN_NODES_PER_AGGREGATE)) {
					int32_t const total = size[r0] + size[rj];

					UNION(r0, rj);
					FIND(r0, r0);
					size[r0] = total;
					changed = true;
				}
			}
		}

		if (!changed) {
			break;
		}
	}

	memset(size, 0, (nn + 1) * sizeof *size);

	for (size_t a = 0; a < nn; a++) {
		int32_t r;
		FIND((int32_t) a, r);
		size[r]++;
	}

	/* largest piece of every bin (root node id; -1: the bin has no piece of three nodes) */

	for (int64_t b = 0; b < n_bins; b++) {
		remap[b] = -1;
	}

	for (size_t a = 0; a < nn; a++) {
		if (parent[a] == (int32_t) a && size[a] >= MIN_NODES_PER_AGGREGATE) {
			int32_t const b = node_agg[a];

			if (remap[b] < 0 || size[a] > size[remap[b]]) {
				remap[b] = (int32_t) a;
			}
		}
	}

	for (size_t a = 0; a < nn; a++) {
		if (parent[a] != (int32_t) a || size[a] >= MIN_NODES_PER_AGGREGATE) {
			continue;
		}

		int64_t const bx = node_agg[a] % nbx;
		int64_t const by = node_agg[a] / nbx;
		int32_t best = remap[node_agg[a]];

		for (int64_t ring = 1; best < 0 && ring < nbx + nby; ring++) {
			for (int64_t dy = -ring; dy <= ring; dy++) {
				for (int64_t dx = -ring; dx <= ring; dx++) {
					int64_t const cx = bx + dx;
					int64_t const cy = by + dy;

					if ((llabs(dx) != ring && llabs(dy) != ring) || cx < 0 || cy < 0 || cx >= nbx || cy >= nby) {
						continue;
					}

					int32_t const cand = remap[cy * nbx + cx];

					if (cand >= 0 && (best < 0 || size[cand] > size[best])) {
						best = cand;
					}
				}
			}
		}

		if (best < 0) { /* no piece of the whole mesh holds three nodes: no coarse level */
			free(parent);
			free(size);
			free(count);
			free(remap);
			return 0;
		}

		parent[a] = best; /* best is a root of a big piece and stays one: big roots are never re-parented here */
	}

	/* compact: ids in order of each piece's smallest node (its root) - deterministic */

	int32_t n_agg = 0;

	for (size_t a = 0; a < nn; a++) {
		size[a] = parent[a] == (int32_t) a ? n_agg++ : -1;
	}

	bool const ok = n_agg > 0;

	for (size_t a = 0; a < nn; a++) {
		int32_t r;
		FIND((int32_t) a, r);
		node_agg[a] = size[r];
	}

#undef FIND
#undef UNION

	free(parent);
	free(size);
	free(count);
	free(remap);

	return ok ? n_agg : 0;
}

bfmi_coarse_t* bfmi_coarse_build(bfm_state_t* state, bfm_mesh_t const* gmesh, bfmi_part_t const* part, int32_t target) {
	size_t const nn = gmesh->n_nodes;
	size_t const kind = gmesh->kind;

	(void) state;

	if (target < 4 || nn < (size_t) target * MIN_NODES_PER_AGGREGATE) {
		return NULL;
	}

	int32_t* const node_agg = malloc((nn + 1) * sizeof *node_agg);

	if (node_agg == NULL) {
		return NULL;
	}

	int32_t const n_agg = aggregate_nodes(gmesh, target, node_agg);

	if (n_agg < 4) {
		free(node_agg);
		return NULL;
	}

	bfmi_coarse_t* const c = calloc(1, sizeof *c);
	double* const cen = calloc((size_t) n_agg * 3, sizeof *cen); /* sum x, sum y, count */
	uint64_t* pairs = NULL;

	if (c == NULL || cen == NULL) {
		goto fail;
	}

	c->n_agg = n_agg;

	/* centroids over the global mesh, fixed order -> identical on every rank */

	for (size_t a = 0; a < nn; a++) {
		double* const s = &cen[3 * node_agg[a]];

		s[0] += gmesh->coords[2 * a + 0];
		s[1] += gmesh->coords[2 * a + 1];
		s[2] += 1;
	}

	/* per local node: aggregate + position relative to its centroid; per aggregate: its OWNED nodes */

	size_t const n_local = part != NULL ? (size_t) part->n_local : nn;
	size_t const own_begin = part != NULL ? (size_t) part->own_begin : 0;
	size_t const own_end = part != NULL ? (size_t) part->own_end : nn;

	c->n_local = (int32_t) n_local;
	c->agg = malloc((n_local + 1) * sizeof *c->agg);
	c->wgeom = malloc((n_local + 1) * 2 * sizeof *c->wgeom);
	c->agg_ptr = calloc((size_t) n_agg + 2, sizeof *c->agg_ptr);
	c->agg_nodes = malloc((own_end - own_begin + 1) * sizeof *c->agg_nodes);

	if (c->agg == NULL || c->wgeom == NULL || c->agg_ptr == NULL || c->agg_nodes == NULL) {
		goto fail;
	}

	for (size_t l = 0; l < n_local; l++) {
		size_t const a = part != NULL ? part->l2g[l] : l;
		int32_t const g = node_agg[a];

		c->agg[l] = g;
		c->wgeom[2 * l + 0] = gmesh->coords[2 * a + 0] - cen[3 * g + 0] / cen[3 * g + 2];
		c->wgeom[2 * l + 1] = gmesh->coords[2 * a + 1] - cen[3 * g + 1] / cen[3 * g + 2];

		if (l >= own_begin && l < own_end) {
			c->agg_ptr[g + 2]++;
		}
	}

	for (int32_t g = 0; g < n_agg; g++) {
		c->agg_ptr[g + 2] += c->agg_ptr[g + 1];
	}

	for (size_t l = own_begin; l < own_end; l++) {
		c->agg_nodes[c->agg_ptr[c->agg[l] + 1]++] = (int32_t) l;
	}

	/* aggregate graph from the global elements: g ~ h when an element has nodes in both */

	size_t n_pairs = 0;
	size_t cap_pairs = 1 << 16;

	pairs = malloc(cap_pairs * sizeof *pairs);

	if (pairs == NULL) {
		goto fail;
	}

	for (size_t e = 0; e < gmesh->n_elems; e++) {
		size_t const* const el = &gmesh->elems[e * kind];
		int32_t const g0 = node_agg[el[0]];
		bool mixed = false;

		for (size_t j = 1; j < kind; j++) {
			mixed |= node_agg[el[j]] != g0;
		}

		if (!mixed) {
			continue;
		}

		if (n_pairs + kind * kind > cap_pairs) {
			/* dedupe before growing: boundary elements repeat the same few pairs */
			qsort(pairs, n_pairs, sizeof *pairs, cmp_u64);

			size_t uniq = 0;

			for (size_t i = 0; i < n_pairs; i++) {
				if (i == 0 || pairs[i] != pairs[i - 1]) {
					pairs[uniq++] = pairs[i];
				}
			}

			n_pairs = uniq;

			if (n_pairs + kind * kind > cap_pairs / 2) {
				cap_pairs *= 2;
				uint64_t* const grown = realloc(pairs, cap_pairs * sizeof *pairs);

				if (grown == NULL) {
					goto fail;
				}

				pairs = grown;
			}
		}

		for (size_t j = 0; j < kind; j++) {
			for (size_t k = 0; k < kind; k++) {
				int32_t const g = node_agg[el[j]];
				int32_t const h = node_agg[el[k]];

				if (g != h) {
					pairs[n_pairs++] = (uint64_t) g << 32 | (uint32_t) h;
				}
			}
		}
	}

	qsort(pairs, n_pairs, sizeof *pairs, cmp_u64);

	{
		size_t uniq = 0;

		for (size_t i = 0; i < n_pairs; i++) {
			if (i == 0 || pairs[i] != pairs[i - 1]) {
				pairs[uniq++] = pairs[i];
			}
		}

		n_pairs = uniq;
	}

	for (size_t i = 0; i < n_pairs; i++) {
		int64_t const span = llabs((int64_t) (pairs[i] >> 32) - (int64_t) (pairs[i] & 0xffffffffu));
		c->agg_span = span > c->agg_span ? (int32_t) span : c->agg_span;
	}

	/* CSR of the aggregate graph (both directions are present: the pair loop above is symmetric) */

	int32_t* const adj_ptr = calloc((size_t) n_agg + 2, sizeof *adj_ptr);
	int32_t* const adj = malloc((n_pairs + 1) * sizeof *adj);
	int32_t* const mark = malloc(((size_t) n_agg + 1) * sizeof *mark);

	c->color = malloc(((size_t) n_agg + 1) * sizeof *c->color);

	if (adj_ptr == NULL || adj == NULL || mark == NULL || c->color == NULL) {
		free(adj_ptr);
		free(adj);
		free(mark);
		goto fail;
	}

	for (size_t i = 0; i < n_pairs; i++) {
		adj_ptr[(pairs[i] >> 32) + 1]++;
	}

	for (int32_t g = 0; g < n_agg; g++) {
		adj_ptr[g + 1] += adj_ptr[g];
	}

	for (size_t i = 0; i < n_pairs; i++) { /* sorted by g then h: fills in order */
		adj[i] = (int32_t) (pairs[i] & 0xffffffffu);
	}

	/* greedy distance-2 colouring: g differs from every aggregate within two hops */

	int32_t n_colors = 0;

	for (int32_t g = 0; g < n_agg; g++) {
		mark[g] = -1;
		c->color[g] = -1;
	}

	for (int32_t g = 0; g < n_agg; g++) {
		/* mark[colour] = g for every colour taken within distance 2 (mark is indexed by colour; there are
		 * never more colours than aggregates) */

		for (int32_t t = adj_ptr[g]; t < adj_ptr[g + 1]; t++) {
			int32_t const h = adj[t];

			if (c->color[h] >= 0) {
				mark[c->color[h]] = g;
			}

			for (int32_t u = adj_ptr[h]; u < adj_ptr[h + 1]; u++) {
				int32_t const k = adj[u];

				if (k != g && c->color[k] >= 0) {
					mark[c->color[k]] = g;
				}
			}
		}

		int32_t col = 0;

		while (mark[col] == g) {
			col++;
		}

		c->color[g] = col;
		n_colors = col + 1 > n_colors ? col + 1 : n_colors;
	}

	c->n_colors = n_colors;
	c->color_nbr = malloc((size_t) n_agg * (size_t) n_colors * sizeof *c->color_nbr);

	if (c->color_nbr == NULL) {
		free(adj_ptr);
		free(adj);
		free(mark);
		goto fail;
	}

	for (size_t i = 0; i < (size_t) n_agg * (size_t) n_colors; i++) {
		c->color_nbr[i] = -1;
	}

	for (int32_t g = 0; g < n_agg; g++) {
		c->color_nbr[(size_t) g * n_colors + c->color[g]] = g;

		for (int32_t t = adj_ptr[g]; t < adj_ptr[g + 1]; t++) {
			c->color_nbr[(size_t) g * n_colors + c->color[adj[t]]] = adj[t]; /* unique per colour by construction */
		}
	}

	free(adj_ptr);
	free(adj);
	free(mark);
	free(pairs);
	free(cen);
	free(node_agg);

	return c;

fail:

	free(pairs);
	free(cen);
	free(node_agg);
	bfmi_coarse_free(c);

	return NULL;
}

static int mirror(void** d_ptr, void const* src, size_t bytes) {
	if (bfmg_alloc(d_ptr, bytes) < 0) {
		return -1;
	}

	return bfmg_upload(*d_ptr, src, bytes);
}

int bfmi_coarse_upload(bfmi_coarse_t* c) {
	bfmg_coarse_t* const d = &c->dev;

	if (c->on_device) {
		return 0;
	}

	c->on_device = true; /* bfmi_coarse_free releases whatever got allocated */

	d->n_agg = c->n_agg;
	d->n_colors = c->n_colors;
	d->nc = (3 * c->n_agg + 31) / 32 * 32;
	d->half_bw = 3 * c->agg_span + 2;

	if (
		mirror((void**) &d->agg, c->agg, (size_t) c->n_local * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->wgeom, c->wgeom, (size_t) c->n_local * 2 * sizeof(double)) < 0 ||
		mirror((void**) &d->agg_ptr, c->agg_ptr, ((size_t) c->n_agg + 1) * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->agg_nodes, c->agg_nodes, ((size_t) c->agg_ptr[c->n_agg] + 1) * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->color_nbr, c->color_nbr, (size_t) c->n_agg * (size_t) c->n_colors * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->color, c->color, (size_t) c->n_agg * sizeof(int32_t)) < 0
	) {
		return -1;
	}

	return 0;
}

/* ---- cache --------------------------------------------------------------------------------------------
 * examples/benchmark.py-style loops call bfm_sim_run again and again on one mesh: aggregates, colouring
 * and their device mirrors are kept as long as mesh identity, connectivity, coordinates and partition
 * are the same. */

static bfmi_coarse_t* cached;
static bool cached_none;          /* the last key produced no coarse level */
static bfmi_coarse_t cached_key;  /* key of that negative result */

static uint64_t hash_coords(bfm_mesh_t const* mesh) {
	size_t const count = mesh->n_nodes * 2;
	size_t const chunk = 1 << 16;
	size_t const n_chunks = (count + chunk - 1) / chunk;
	uint64_t total = 0x51ed270b7a2c3f11ull ^ count;

#pragma omp parallel for schedule(static) reduction(^ : total) if (n_chunks > 16)
	for (size_t c = 0; c < n_chunks; c++) {
		size_t const end = (c + 1) * chunk < count ? (c + 1) * chunk : count;
		uint64_t h = 0xcbf29ce484222325ull + c;

		for (size_t i = c * chunk; i < end; i++) {
			uint64_t bits;
			memcpy(&bits, &mesh->coords[i], sizeof bits);
			h = (h ^ bits) * 0x100000001b3ull;
			h ^= h >> 31;
		}

		total ^= h * (2 * c + 1);
	}

	return total;
}

static bool same_key(bfmi_coarse_t const* c, bfm_mesh_t const* gmesh, uint64_t elems_hash, uint64_t coords_hash, int rank, int world, int32_t target) {
	return c->gmesh == gmesh && c->key_nodes == gmesh->n_nodes && c->key_elems == gmesh->n_elems && c->elems_hash == elems_hash && c->coords_hash == coords_hash && c->rank == rank && c->world == world && c->target == target;
}

static void set_key(bfmi_coarse_t* c, bfm_mesh_t const* gmesh, uint64_t elems_hash, uint64_t coords_hash, int rank, int world, int32_t target) {
	c->gmesh = gmesh;
	c->key_nodes = gmesh->n_nodes;
	c->key_elems = gmesh->n_elems;
	c->elems_hash = elems_hash;
	c->coords_hash = coords_hash;
	c->rank = rank;
	c->world = world;
	c->target = target;
}

void bfmi_coarse_release(bfmi_coarse_t* c) {
	if (c != NULL && __atomic_sub_fetch(&c->refs, 1, __ATOMIC_ACQ_REL) == 0) {
		bfmi_coarse_free(c);
	}
}

bfmi_coarse_t* bfmi_coarse_for_mesh(bfm_state_t* state, bfm_mesh_t const* gmesh, uint64_t elems_hash, bfmi_part_t const* part, int32_t target, bool* none) {
	int const rank = part != NULL ? part->rank : 0;
	int const world = part != NULL ? part->world : 1;
	uint64_t const coords_hash = hash_coords(gmesh);

	*none = false;

	if (cached != NULL && same_key(cached, gmesh, elems_hash, coords_hash, rank, world, target)) {
		__atomic_add_fetch(&cached->refs, 1, __ATOMIC_RELAXED);
		return cached;
	}

	if (cached_none && same_key(&cached_key, gmesh, elems_hash, coords_hash, rank, world, target)) {
		*none = true;
		return NULL;
	}

	bfmi_coarse_t* const c = bfmi_coarse_build(state, gmesh, part, target);

	if (c == NULL) {
		cached_none = true;
		set_key(&cached_key, gmesh, elems_hash, coords_hash, rank, world, target);
		*none = true;
		return NULL;
	}

	c->refs = 1; /* the cache's */
	set_key(c, gmesh, elems_hash, coords_hash, rank, world, target);

	bfmi_coarse_release(cached);
	cached = c;

	__atomic_add_fetch(&c->refs, 1, __ATOMIC_RELAXED); /* the caller's */
	return c;
}

void bfmi_coarse_forget(bfm_mesh_t const* gmesh) {
	if (cached != NULL && cached->gmesh == gmesh) {
		bfmi_coarse_release(cached);
		cached = NULL;
	}

	if (cached_none && cached_key.gmesh == gmesh) {
		cached_none = false;
	}
}

/* introspection for tests (bfm_b200.h) */
int bfmx_coarse_plan(bfm_mesh_t* mesh, int target, int32_t* n_aggregates, int32_t* n_colors, int32_t* node_aggregate, int32_t* aggregate_color) {
	bfmi_coarse_t* const c = bfmi_coarse_build(mesh->state, mesh, NULL, target);

	if (c == NULL) {
		*n_aggregates = 0;
		*n_colors = 0;
		return 0;
	}

	*n_aggregates = c->n_agg;
	*n_colors = c->n_colors;

	if (node_aggregate != NULL) {
		memcpy(node_aggregate, c->agg, mesh->n_nodes * sizeof *node_aggregate);
	}

	if (aggregate_color != NULL) {
		memcpy(aggregate_color, c->color, (size_t) c->n_agg * sizeof *aggregate_color);
	}

	bfmi_coarse_free(c);
	return 0;
}
