/*
 * plan.c - symbolic phase: from mesh connectivity to the SELL-32 node-block sparsity pattern and
 * the deterministic element-to-nonzero ("contributor") map the assembly kernel walks.
 *
 * The pattern is the STRUCTURAL one - block (a, b) exists iff nodes a and b share an element - which
 * is exactly the set of entries the reference's dense assembly ever touches (system.c:207-223);
 * which of those end up numerically zero is decided by the arithmetic, on the GPU.  For each block
 * the contributor list holds (element, local row node j, local column node k), sorted by element,
 * then j, then k: the order in which the reference's element loop (system.c:460) adds to that entry.
 *
 * Built once per mesh and cached by mesh identity + a hash of the connectivity, so repeated bfm_sim_run calls on the
 * same mesh (examples/benchmark.py) reuse it.  With a CUDA device the builder is symbolic.cu (count / scan / per-node
 * sort kernels; the host only narrows the connectivity to 32 bits and receives the arrays host code reads: columns,
 * row lengths, slice offsets, diagonal slots); build_from_mesh below is the same construction on the host (OpenMP
 * over nodes) - what introspection gets on a machine without a device, what BFM_PLAN=host forces, and what the GPU
 * test compares the kernels with, array for array.
 */
#include "internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SLICE 32

/* ---- helpers --------------------------------------------------------------------------------- */

static int cmp_u64(void const* a, void const* b) {
	uint64_t const x = *(uint64_t const*) a;
	uint64_t const y = *(uint64_t const*) b;
	return x < y ? -1 : x > y;
}

static void sort_u64(uint64_t* v, size_t n) {
	if (n > 48) {
		qsort(v, n, sizeof *v, cmp_u64); /* keys are unique, stability is moot */
		return;
	}

	for (size_t i = 1; i < n; i++) {
		uint64_t const cur = v[i];
		size_t j = i;

		for (; j > 0 && v[j - 1] > cur; j--) {
			v[j] = v[j - 1];
		}

		v[j] = cur;
	}
}

/* order-sensitive 64-bit hash of the connectivity, chunked so that OpenMP can help on big meshes.  It runs on every
 * bfm_sim_run (the cached plan is only valid while the connectivity is what it was), over 1.2 GB at 50 M DOF, so the
 * inner loop keeps four independent multiply chains going: memory speed instead of multiplier latency. */
static uint64_t hash_elems_part(size_t const* elems, size_t count, size_t first_chunk, size_t chunk_stride) {
	size_t const chunk = 1 << 16;
	size_t const n_chunks = (count + chunk - 1) / chunk;
	uint64_t total = first_chunk == 0 ? 0x9e3779b97f4a7c15ull ^ count : 0; /* the parts of all ranks XOR to the whole */

#pragma omp parallel for schedule(static) reduction(^ : total) if (n_chunks > 16 * chunk_stride)
	for (size_t c = first_chunk; c < n_chunks; c += chunk_stride) {
		size_t const end = (c + 1) * chunk < count ? (c + 1) * chunk : count;
		uint64_t h0 = 0xcbf29ce484222325ull + c, h1 = 0x84222325cbf29ce4ull ^ c, h2 = 0x9e3779b97f4a7c15ull + 3 * c, h3 = 0xc2b2ae3d27d4eb4full ^ (c << 7);
		size_t i = c * chunk;

		for (; i + 4 <= end; i += 4) {
			h0 = (h0 ^ elems[i + 0]) * 0x100000001b3ull;
			h1 = (h1 ^ elems[i + 1]) * 0x9e3779b97f4a7c15ull;
			h2 = (h2 ^ elems[i + 2]) * 0xc2b2ae3d27d4eb4full;
			h3 = (h3 ^ elems[i + 3]) * 0x165667b19e3779f9ull;
			h0 ^= h0 >> 29, h1 ^= h1 >> 31, h2 ^= h2 >> 27, h3 ^= h3 >> 33;
		}

		for (; i < end; i++) {
			h0 = (h0 ^ elems[i]) * 0x100000001b3ull;
			h0 ^= h0 >> 29;
		}

		uint64_t h = (h0 * 31 + h1) * 0x100000001b3ull;

		h = ((h ^ (h >> 32)) * 31 + h2) * 0x9e3779b97f4a7c15ull;
		h = ((h ^ (h >> 29)) * 31 + h3) * 0xc2b2ae3d27d4eb4full;

		total ^= (h ^ (h >> 31)) * (2 * c + 1);
	}

	return total;
}

static uint64_t hash_elems(size_t const* elems, size_t count) {
	return hash_elems_part(elems, count, 0, 1);
}

uint64_t bfmi_mesh_hash(bfm_mesh_t const* mesh) {
	return hash_elems(mesh->elems, mesh->n_elems * mesh->kind);
}

/* several ranks holding the same mesh (one process per GPU, SPMD): rank r hashes every world-th chunk; the XOR of all
 * parts is bfmi_mesh_hash - each process reads 1 / world of the connectivity instead of all of it */
uint64_t bfmi_mesh_hash_part(bfm_mesh_t const* mesh, int rank, int world) {
	return hash_elems_part(mesh->elems, mesh->n_elems * mesh->kind, (size_t) rank, (size_t) world);
}

static void plan_free(bfmi_plan_t* plan) {
	if (plan->on_device) {
		bfmg_free(plan->dev.slice_off);
		bfmg_free(plan->dev.row_len);
		bfmg_free(plan->dev.scol);
		bfmg_free(plan->dev.diag_pos);
		bfmg_free(plan->dev.ctr_ptr);
		bfmg_free(plan->dev.ctr);
		bfmg_free(plan->dev.elems);
	}

	free(plan->slice_off);
	free(plan->row_len);
	free(plan->scol);
	free(plan->diag_pos);
	free(plan->ctr_ptr);
	free(plan->ctr);
	free(plan->elems32);
	free(plan);
}

void bfmi_plan_retain(bfmi_plan_t* plan) {
	__atomic_add_fetch(&plan->refs, 1, __ATOMIC_RELAXED);
}

void bfmi_plan_release(bfmi_plan_t* plan) {
	if (plan != NULL && __atomic_sub_fetch(&plan->refs, 1, __ATOMIC_ACQ_REL) == 0) {
		plan_free(plan);
	}
}

/* slice offsets + padding columns from row_len; allocates slice_off, scol, diag_pos */
static int layout_slices(bfmi_plan_t* plan) {
	int32_t const nb = plan->nb;

	plan->n_slices = (nb + SLICE - 1) / SLICE;
	plan->slice_off = malloc(((size_t) plan->n_slices + 1) * sizeof *plan->slice_off);

	if (plan->slice_off == NULL) {
		return -1;
	}

	int64_t off = 0;
	int64_t blocks = 0;

	for (int32_t s = 0; s < plan->n_slices; s++) {
		int32_t longest = 0;

		for (int32_t a = s * SLICE; a < nb && a < (s + 1) * SLICE; a++) {
			longest = plan->row_len[a] > longest ? plan->row_len[a] : longest;
			blocks += plan->row_len[a];
		}

		plan->slice_off[s] = (int32_t) off;
		off += (int64_t) longest * SLICE;

		if (off > INT32_MAX) {
			return -1;
		}
	}

	plan->slice_off[plan->n_slices] = (int32_t) off;
	plan->n_slots = off;
	plan->n_blocks = blocks;

	plan->scol = malloc(((size_t) off + 1) * sizeof *plan->scol);
	plan->diag_pos = malloc(((size_t) nb + 1) * sizeof *plan->diag_pos);

	if (plan->scol == NULL || plan->diag_pos == NULL) {
		return -1;
	}

	/* every slot starts as padding: it points at its own row (clamped), carries zeros */

#pragma omp parallel for schedule(static) if (plan->n_slices > 4096)
	for (int32_t s = 0; s < plan->n_slices; s++) {
		for (int32_t slot = plan->slice_off[s]; slot < plan->slice_off[s + 1]; slot++) {
			int32_t const a = s * SLICE + (slot - plan->slice_off[s]) % SLICE;
			plan->scol[slot] = a < nb ? a : nb - 1;
		}
	}

	return 0;
}

static inline int64_t slot_of(bfmi_plan_t const* plan, int32_t a, int32_t t) {
	return (int64_t) plan->slice_off[a / SLICE] + (int64_t) t * SLICE + a % SLICE;
}

int64_t bfmi_plan_find(bfmi_plan_t const* plan, int32_t a, int32_t b) {
	int32_t lo = 0;
	int32_t hi = plan->row_len[a];

	while (lo < hi) { /* columns ascend along a row */
		int32_t const mid = (lo + hi) / 2;

		if (plan->scol[slot_of(plan, a, mid)] < b) {
			lo = mid + 1;
		}

		else {
			hi = mid;
		}
	}

	if (lo < plan->row_len[a] && plan->scol[slot_of(plan, a, lo)] == b) {
		return slot_of(plan, a, lo);
	}

	return -1;
}

/* ---- plan from a mesh ------------------------------------------------------------------------- */

static double plan_now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* the same plan built by symbolic.cu.  The contributor map (ctr_ptr, ctr) stays on the device - only the assembly
 * kernel reads it; bfmx_mesh_pattern_copy fetches it on demand */
static bfmi_plan_t* build_on_device(bfm_mesh_t const* mesh, uint64_t hash) {
	size_t const kind = mesh->kind;
	size_t const nn = mesh->n_nodes;
	size_t const ne = mesh->n_elems;

	if (nn == 0 || nn >= (1u << 30) || ne >= (1u << 28) || ne * kind * kind >= INT32_MAX) {
		return NULL;
	}

	bfmi_plan_t* const plan = calloc(1, sizeof *plan);

	if (plan == NULL) {
		return NULL;
	}

	plan->refs = 1;
	plan->mesh = mesh;
	plan->n_nodes = nn;
	plan->n_elems = ne;
	plan->kind = (int) kind;
	plan->elems_hash = hash;
	plan->nb = (int32_t) nn;

	bool const verbose = getenv("BFM_JOB_VERBOSE") != NULL;
	double t_mark = plan_now_ms();

#define MARK(what) do { if (verbose) { double const now_ = plan_now_ms(); fprintf(stderr, "[plan] %-28s %8.2f ms\n", (what), now_ - t_mark); t_mark = now_; } } while (0)

	int32_t* d_elems = NULL;
	int bad = 0;

	/* connectivity: narrowed to 32 bits on its way into the staging buffers; a node number outside the table fails */
	if (bfmg_alloc((void**) &d_elems, (ne * kind + 1) * sizeof *d_elems) < 0 || bfmg_upload_narrow(d_elems, mesh->elems, ne * kind, nn, &bad) < 0 || bad) {
		bfmg_free(d_elems);
		free(plan);
		return NULL;
	}

	MARK("connectivity to the device");

	if (bfmg_plan_build(plan->nb, (int64_t) ne, (int32_t) kind, d_elems, &plan->dev, &plan->n_ctr) < 0) {
		bfmg_free(d_elems);
		free(plan);
		return NULL;
	}

	MARK("kernels (symbolic.cu)");

	plan->dev.elems = d_elems;
	plan->on_device = true; /* from here on plan_free releases the device arrays */
	plan->built_on_device = true;
	plan->n_slices = plan->dev.n_slices;
	plan->n_slots = plan->dev.n_slots;

	plan->slice_off = malloc(((size_t) plan->n_slices + 1) * sizeof *plan->slice_off);
	plan->row_len = malloc((nn + 1) * sizeof *plan->row_len);
	plan->scol = malloc(((size_t) plan->n_slots + 1) * sizeof *plan->scol);
	plan->diag_pos = malloc((nn + 1) * sizeof *plan->diag_pos);

	if (
		plan->slice_off == NULL || plan->row_len == NULL || plan->scol == NULL || plan->diag_pos == NULL ||
		bfmg_download(plan->slice_off, plan->dev.slice_off, ((size_t) plan->n_slices + 1) * sizeof *plan->slice_off) < 0 ||
		bfmg_download(plan->row_len, plan->dev.row_len, nn * sizeof *plan->row_len) < 0 ||
		bfmg_download(plan->scol, plan->dev.scol, (size_t) plan->n_slots * sizeof *plan->scol) < 0 ||
		bfmg_download(plan->diag_pos, plan->dev.diag_pos, nn * sizeof *plan->diag_pos) < 0
	) {
		plan_free(plan);
		return NULL;
	}

	MARK("pattern to the host");
#undef MARK

	int64_t blocks = 0;

#pragma omp parallel for schedule(static) reduction(+ : blocks) if (nn > 50000)
	for (size_t a = 0; a < nn; a++) {
		blocks += plan->row_len[a];
	}

	plan->n_blocks = blocks;
	plan->h2d_bytes = ne * kind * sizeof(int32_t);
	plan->d2h_bytes = ((size_t) plan->n_slices + 1 + 2 * nn + (size_t) plan->n_slots) * sizeof(int32_t);

	return plan;
}

static bfmi_plan_t* build_from_mesh(bfm_mesh_t const* mesh, uint64_t hash) {
	size_t const kind = mesh->kind;
	size_t const nn = mesh->n_nodes;
	size_t const ne = mesh->n_elems;

	if (nn == 0 || nn >= (1u << 30) || ne >= (1u << 28) || ne * kind * kind >= INT32_MAX) {
		return NULL;
	}

	bfmi_plan_t* const plan = calloc(1, sizeof *plan);

	if (plan == NULL) {
		return NULL;
	}

	plan->refs = 1;
	plan->mesh = mesh;
	plan->n_nodes = nn;
	plan->n_elems = ne;
	plan->kind = (int) kind;
	plan->elems_hash = hash;
	plan->nb = (int32_t) nn;

	/* node -> (element, local index) incidence, elements ascending */

	int64_t* inc_off = calloc(nn + 2, sizeof *inc_off);
	uint32_t* inc = malloc((ne * kind + 1) * sizeof *inc);
	uint64_t* keys = malloc((ne * kind * kind + 1) * sizeof *keys);

	plan->elems32 = malloc((ne * kind + 1) * sizeof *plan->elems32);
	plan->row_len = malloc((nn + 1) * sizeof *plan->row_len);

	bool ok = inc_off != NULL && inc != NULL && keys != NULL && plan->elems32 != NULL && plan->row_len != NULL;

	for (size_t i = 0; ok && i < ne * kind; i++) {
		if (mesh->elems[i] >= nn) {
			ok = false; /* connectivity points outside the node table */
			break;
		}

		plan->elems32[i] = (int32_t) mesh->elems[i];
		inc_off[mesh->elems[i] + 2]++;
	}

	if (!ok) {
		goto fail;
	}

	for (size_t a = 0; a < nn; a++) {
		inc_off[a + 2] += inc_off[a + 1];
	}

	for (size_t e = 0; e < ne; e++) {
		for (size_t j = 0; j < kind; j++) {
			inc[inc_off[mesh->elems[e * kind + j] + 1]++] = (uint32_t) (e << 2 | j);
		}
	}

	/* per node: one key per (incident element, its local index j, local column k), ordered by
	 * (column node, element, j, k); row length = distinct column nodes (the diagonal always exists) */

#pragma omp parallel for schedule(dynamic, 1024) if (nn > 50000)
	for (size_t a = 0; a < nn; a++) {
		uint64_t* const seg = &keys[inc_off[a] * kind];
		size_t cnt = 0;

		for (int64_t t = inc_off[a]; t < inc_off[a + 1]; t++) {
			uint32_t const e = inc[t] >> 2;
			uint32_t const j = inc[t] & 3;

			for (uint32_t k = 0; k < kind; k++) {
				uint64_t const b = (uint64_t) plan->elems32[e * kind + k];
				seg[cnt++] = b << 32 | (uint64_t) e << 4 | j << 2 | k;
			}
		}

		sort_u64(seg, cnt);

		int32_t len = 0;

		for (size_t t = 0; t < cnt; t++) {
			len += t == 0 || seg[t] >> 32 != seg[t - 1] >> 32;
		}

		plan->row_len[a] = cnt ? len : 1;
	}

	if (layout_slices(plan) < 0) {
		goto fail;
	}

	plan->ctr_ptr = calloc((size_t) plan->n_slots + 2, sizeof *plan->ctr_ptr);
	plan->ctr = malloc((ne * kind * kind + 1) * sizeof *plan->ctr);

	if (plan->ctr_ptr == NULL || plan->ctr == NULL) {
		goto fail;
	}

	/* pass 1: columns and per-slot contribution counts (stored one ahead for the prefix sum) */

#pragma omp parallel for schedule(dynamic, 1024) if (nn > 50000)
	for (size_t a = 0; a < nn; a++) {
		uint64_t const* const seg = &keys[inc_off[a] * kind];
		size_t const cnt = (size_t) (inc_off[a + 1] - inc_off[a]) * kind;

		int32_t t = -1;

		/* a node inside an element always meets itself (k == j), so the diagonal block is among the
		 * keys; a node in no element gets a lone, empty diagonal block in slot 0 (already labelled
		 * with its own row by layout_slices) */

		plan->diag_pos[a] = (int32_t) slot_of(plan, (int32_t) a, 0);

		for (size_t i = 0; i < cnt; i++) {
			uint64_t const b = seg[i] >> 32;

			if (i == 0 || b != seg[i - 1] >> 32) {
				t++;
				plan->scol[slot_of(plan, (int32_t) a, t)] = (int32_t) b;

				if (b == a) {
					plan->diag_pos[a] = (int32_t) slot_of(plan, (int32_t) a, t);
				}
			}

			plan->ctr_ptr[slot_of(plan, (int32_t) a, t) + 1]++;
		}
	}

	for (int64_t s = 0; s < plan->n_slots; s++) {
		plan->ctr_ptr[s + 1] += plan->ctr_ptr[s];
	}

	plan->n_ctr = plan->ctr_ptr[plan->n_slots];

	/* pass 2: the packed (element, j, k) lists, slot by slot */

#pragma omp parallel for schedule(dynamic, 1024) if (nn > 50000)
	for (size_t a = 0; a < nn; a++) {
		uint64_t const* const seg = &keys[inc_off[a] * kind];
		size_t const cnt = (size_t) (inc_off[a + 1] - inc_off[a]) * kind;

		int32_t t = -1;
		int32_t fill = 0;

		for (size_t i = 0; i < cnt; i++) {
			if (i == 0 || seg[i] >> 32 != seg[i - 1] >> 32) {
				fill = plan->ctr_ptr[slot_of(plan, (int32_t) a, ++t)];
			}

			plan->ctr[fill++] = (uint32_t) seg[i];
		}
	}

	free(inc_off);
	free(inc);
	free(keys);

	return plan;

fail:

	free(inc_off);
	free(inc);
	free(keys);
	plan_free(plan);

	return NULL;
}

/* ---- cache ------------------------------------------------------------------------------------ */

#define CACHE_SLOTS 8

static struct {
	bfmi_plan_t* plan;
	uint64_t stamp;
} cache[CACHE_SLOTS];

static uint64_t cache_clock;

bfmi_plan_t* bfmi_plan_for_mesh(bfm_state_t* state, bfm_mesh_t const* mesh) {
	return bfmi_plan_for_mesh_hashed(state, mesh, hash_elems(mesh->elems, mesh->n_elems * mesh->kind));
}

/* hash = bfmi_mesh_hash(mesh), for callers that have it already (1.2 GB to read at 50 M DOF) */
bfmi_plan_t* bfmi_plan_for_mesh_hashed(bfm_state_t* state, bfm_mesh_t const* mesh, uint64_t hash) {
	if (mesh->dim != 2 || (mesh->kind != BFM_ELEM_KIND_SIMPLEX && mesh->kind != BFM_ELEM_KIND_QUAD)) {
		return NULL; /* reference system.c:436-442 */
	}

	int victim = 0;

	for (int i = 0; i < CACHE_SLOTS; i++) {
		bfmi_plan_t* const p = cache[i].plan;

		if (p != NULL && p->mesh == mesh && p->n_nodes == mesh->n_nodes && p->n_elems == mesh->n_elems && p->kind == (int) mesh->kind && p->elems_hash == hash) {
			cache[i].stamp = ++cache_clock;
			bfmi_plan_retain(p);
			return p;
		}

		if (cache[i].stamp < cache[victim].stamp) {
			victim = i;
		}
	}

	/* with a device the kernels of symbolic.cu build it (BFM_PLAN=host: the OpenMP twin above, for comparison) */
	char const* const how = getenv("BFM_PLAN");
	bool const on_device = bfmg_available() && (how == NULL || strcmp(how, "host") != 0);
	bfmi_plan_t* const plan = on_device ? build_on_device(mesh, hash) : build_from_mesh(mesh, hash);

	if (plan == NULL) {
		BFMI_FAIL(state, "cannot build the sparsity plan (mesh too large for 32-bit indices, bad connectivity or out of memory%s%s)", on_device ? "; " : "", on_device ? bfmg_last_error() : "");
		return NULL;
	}

	for (int i = 0; i < CACHE_SLOTS; i++) { /* a stale entry for the same mesh object goes first */
		if (cache[i].plan != NULL && cache[i].plan->mesh == mesh) {
			victim = i;
			break;
		}
	}

	bfmi_plan_release(cache[victim].plan);

	cache[victim].plan = plan; /* the cache's reference */
	cache[victim].stamp = ++cache_clock;

	bfmi_plan_retain(plan);    /* the caller's reference */
	return plan;
}

void bfmi_plan_forget(bfm_mesh_t const* mesh) {
	for (int i = 0; i < CACHE_SLOTS; i++) {
		if (cache[i].plan != NULL && cache[i].plan->mesh == mesh) {
			bfmi_plan_release(cache[i].plan);
			cache[i].plan = NULL;
			cache[i].stamp = 0;
		}
	}
}

/* ---- plan from a scalar CSR pattern (bfmx_matrix_csr_create) ---------------------------------- */

bfmi_plan_t* bfmi_plan_from_csr(bfm_state_t* state, size_t n, size_t const* rowptr, size_t const* col) {
	(void) state;

	if (n == 0 || n % 2 != 0 || n / 2 >= (1u << 30)) {
		return NULL;
	}

	bfmi_plan_t* const plan = calloc(1, sizeof *plan);

	if (plan == NULL) {
		return NULL;
	}

	plan->refs = 1;
	plan->nb = (int32_t) (n / 2);
	plan->row_len = malloc(((size_t) plan->nb + 1) * sizeof *plan->row_len);

	size_t max_row = 1;

	for (size_t a = 0; a < n / 2; a++) {
		size_t const c = rowptr[2 * a + 2] - rowptr[2 * a] + 1;
		max_row = c > max_row ? c : max_row;
	}

	uint64_t* const scratch = malloc(max_row * sizeof *scratch);

	if (plan->row_len == NULL || scratch == NULL) {
		goto fail;
	}

	for (int pass = 0; pass < 2; pass++) {
		for (size_t a = 0; a < n / 2; a++) {
			size_t cnt = 0;

			scratch[cnt++] = a;

			for (size_t t = rowptr[2 * a]; t < rowptr[2 * a + 2]; t++) {
				if (col[t] >= n) {
					goto fail;
				}

				scratch[cnt++] = col[t] / 2;
			}

			sort_u64(scratch, cnt);

			int32_t len = 0;

			for (size_t t = 0; t < cnt; t++) {
				if (t != 0 && scratch[t] == scratch[t - 1]) {
					continue;
				}

				if (pass == 1) {
					int64_t const slot = slot_of(plan, (int32_t) a, len);
					plan->scol[slot] = (int32_t) scratch[t];

					if (scratch[t] == a) {
						plan->diag_pos[a] = (int32_t) slot;
					}
				}

				len++;
			}

			plan->row_len[a] = len;
		}

		if (pass == 0 && layout_slices(plan) < 0) {
			goto fail;
		}
	}

	free(scratch);
	return plan;

fail:

	free(scratch);
	plan_free(plan);

	return NULL;
}

/* ---- device mirror ---------------------------------------------------------------------------- */

static int mirror(void** d_ptr, void const* src, size_t bytes) {
	if (src == NULL) {
		*d_ptr = NULL;
		return 0;
	}

	if (bfmg_alloc(d_ptr, bytes) < 0) {
		return -1;
	}

	return bfmg_upload(*d_ptr, src, bytes);
}

int bfmi_plan_upload(bfm_state_t* state, bfmi_plan_t* plan) {
	if (plan->on_device) {
		return 0;
	}

	if (!bfmg_available()) {
		return BFMI_FAIL(state, "%s", bfmg_last_error());
	}

	bfmg_pattern_t* const d = &plan->dev;

	d->nb = plan->nb;
	d->n_slices = plan->n_slices;
	d->n_slots = plan->n_slots;
	d->kind = plan->kind;
	d->row_lo = 0;
	d->row_hi = plan->nb;

	plan->on_device = true; /* from here on plan_free releases whatever was allocated */
	plan->h2d_bytes = ((size_t) plan->n_slices + 1 + 2 * (size_t) plan->nb + 2 * (size_t) plan->n_slots + 1 + (size_t) plan->n_ctr + plan->n_elems * (size_t) plan->kind) * sizeof(int32_t);

	if (
		mirror((void**) &d->slice_off, plan->slice_off, ((size_t) plan->n_slices + 1) * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->row_len, plan->row_len, (size_t) plan->nb * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->scol, plan->scol, (size_t) plan->n_slots * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->diag_pos, plan->diag_pos, (size_t) plan->nb * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->ctr_ptr, plan->ctr_ptr, ((size_t) plan->n_slots + 1) * sizeof(int32_t)) < 0 ||
		mirror((void**) &d->ctr, plan->ctr, (size_t) plan->n_ctr * sizeof(uint32_t)) < 0 ||
		mirror((void**) &d->elems, plan->elems32, plan->n_elems * (size_t) plan->kind * sizeof(int32_t)) < 0 ||
		bfmg_sync() < 0
	) {
		return BFMI_FAIL(state, "uploading the sparsity plan failed: %s", bfmg_last_error());
	}

	return 0;
}

/* ---- introspection (bfm_b200.h) ------------------------------------------------------------------ */

int bfmx_mesh_pattern_sizes(bfm_mesh_t* mesh, size_t* n_slices, size_t* n_slots, size_t* n_blocks, size_t* n_contributions) {
	bfmi_plan_t* const plan = bfmi_plan_for_mesh(mesh->state, mesh);

	if (plan == NULL) {
		return -1;
	}

	*n_slices = (size_t) plan->n_slices;
	*n_slots = (size_t) plan->n_slots;
	*n_blocks = (size_t) plan->n_blocks;
	*n_contributions = (size_t) plan->n_ctr;

	bfmi_plan_release(plan);
	return 0;
}

int bfmx_mesh_pattern_copy(bfm_mesh_t* mesh, int32_t* slice_off, int32_t* row_len, int32_t* scol, int32_t* diag_pos, int32_t* ctr_ptr, uint32_t* ctr) {
	bfmi_plan_t* const plan = bfmi_plan_for_mesh(mesh->state, mesh);

	if (plan == NULL) {
		return -1;
	}

	int rv = 0;

	memcpy(slice_off, plan->slice_off, ((size_t) plan->n_slices + 1) * sizeof *slice_off);
	memcpy(row_len, plan->row_len, (size_t) plan->nb * sizeof *row_len);
	memcpy(scol, plan->scol, (size_t) plan->n_slots * sizeof *scol);
	memcpy(diag_pos, plan->diag_pos, (size_t) plan->nb * sizeof *diag_pos);

	if (plan->ctr_ptr != NULL) {
		memcpy(ctr_ptr, plan->ctr_ptr, ((size_t) plan->n_slots + 1) * sizeof *ctr_ptr);
		memcpy(ctr, plan->ctr, (size_t) plan->n_ctr * sizeof *ctr);
	}

	else { /* built by symbolic.cu: the contributor map lives on the device only */
		rv = bfmg_download(ctr_ptr, plan->dev.ctr_ptr, ((size_t) plan->n_slots + 1) * sizeof *ctr_ptr) < 0 || bfmg_download(ctr, plan->dev.ctr, (size_t) plan->n_ctr * sizeof *ctr) < 0 ? -1 : 0;
	}

	bfmi_plan_release(plan);
	return rv;
}
