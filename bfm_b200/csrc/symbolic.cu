/*
 * symbolic.cu - the once-per-mesh symbolic phase on the device (sm_100a): the SELL-32 node-block sparsity pattern
 * with its deterministic element-to-nonzero map (what plan.c's host builder produces, array for array), and the
 * edge list of a mesh in the reference's order (compute_edges, reference mesh.c:32-102).
 *
 * Both are "group by a node, order inside the group" problems, and both orders are TOTAL orders on unique keys, so
 * they are built the same way - no global sort, no floating point, integer atomics only where the result cannot
 * depend on their order:
 *
 *   1. count the records of every node            (integer atomicAdd: a count is order-independent)
 *   2. exclusive scan of the counts               (three kernels: tile sums, one CTA over the tiles, apply)
 *   3. drop each record into its node's segment   (atomic cursor: WHERE in the segment is arbitrary ...)
 *   4. one thread per node sorts its segment      (... and does not matter: the keys are unique, the sorted
 *                                                  segment is the same whatever order the records arrived in)
 *   5. one thread per node walks its sorted segment and writes the output (columns / contributor lists / edges)
 *
 * Sparsity pattern: a record is (row node a; column node b, element e, local row j, local column k), key
 * b << 32 | e << 4 | j << 2 | k - the contributor order of plan.c (element, then j, then k: the order in which the
 * reference's element loop, system.c:460, adds to an entry).  Segments are ~21 keys on a triangulated plate, 16 on
 * quads: each thread sorts its keys in local memory (interleaved per lane, hence coalesced).
 *
 * Edges: a record is a half-edge h = element * sides + side, grouped by its smaller node, key larger node << 32 | h.
 * The reference sorts by (smaller node DESCENDING, larger node ascending) with glibc's stable merge sort - i.e.
 * ties keep ascending h - and then fuses neighbouring opposite half-edges greedily; a fusion can only happen between
 * two half-edges of the same node pair, so the greedy walk never leaves a node's segment and one thread per node
 * replays it exactly.  The very last half-edge of the sorted list is emitted only as somebody's partner (the
 * reference's loop stops one short), which the walk of the last non-empty segment reproduces.
 *
 * Everything runs on the library stream; the caller owns the returned device arrays (bfmg_free).
 */
#include "gpu_internal.cuh"

namespace {

constexpr int kScanItems = 8;
constexpr int kScanTile = kBlock * kScanItems;

/* exclusive prefix of v over the CTA (kBlock threads), total in `total`; two barriers */
__device__ __forceinline__ int64_t block_scan(int64_t v, int64_t& total) {
	__shared__ int64_t warp_total[kWarpsPerBlock];

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = threadIdx.x / kWarp;
	int64_t inc = v;

#pragma unroll
	for (int off = 1; off < kWarp; off <<= 1) {
		int64_t const t = __shfl_up_sync(0xffffffffu, inc, off);
		inc += lane >= off ? t : 0;
	}

	if (lane == kWarp - 1) {
		warp_total[warp] = inc;
	}

	__syncthreads();

	int64_t base = 0;
	int64_t sum = 0;

#pragma unroll
	for (int w = 0; w < kWarpsPerBlock; w++) {
		int64_t const t = warp_total[w];
		base += w < warp ? t : 0;
		sum += t;
	}

	__syncthreads(); /* warp_total may be rewritten by the next call */

	total = sum;
	return base + inc - v;
}

/* ---- exclusive scan of int32 counts (n + 1 entries out: out[n] = total) ---------------------------------- */

__global__ void __launch_bounds__(kBlock) k_scan_tile_sums(int32_t const* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sum) {
	pdl_sync();

	int64_t const first = (int64_t) blockIdx.x * kScanTile + (int64_t) threadIdx.x * kScanItems;
	int64_t mine = 0;

#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		mine += first + k < n ? in[first + k] : 0;
	}

	int64_t total;
	block_scan(mine, total);

	if (threadIdx.x == 0) {
		tile_sum[blockIdx.x] = total;
	}
}

/* one CTA: tile_sum becomes its own exclusive prefix, tile_sum[n_tiles] the grand total */
__global__ void __launch_bounds__(kBlock) k_scan_tiles(int64_t* __restrict__ tile_sum, int64_t n_tiles) {
	pdl_sync();

	int64_t carry = 0;

	for (int64_t first = 0; first < n_tiles; first += kBlock) {
		int64_t const i = first + threadIdx.x;
		int64_t const v = i < n_tiles ? tile_sum[i] : 0;
		int64_t total;
		int64_t const before = block_scan(v, total);

		if (i < n_tiles) {
			tile_sum[i] = carry + before;
		}

		carry += total;
	}

	if (threadIdx.x == 0) {
		tile_sum[n_tiles] = carry;
	}
}

__global__ void __launch_bounds__(kBlock) k_scan_apply(int32_t const* in, int32_t* out, int64_t n, int64_t const* __restrict__ tile_sum, int64_t n_tiles) {
	pdl_sync();

	int64_t const first = (int64_t) blockIdx.x * kScanTile + (int64_t) threadIdx.x * kScanItems;
	int32_t v[kScanItems];
	int64_t mine = 0;

#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		v[k] = first + k < n ? in[first + k] : 0;
		mine += v[k];
	}

	int64_t total;
	int64_t at = tile_sum[blockIdx.x] + block_scan(mine, total); /* in == out is fine: every thread has read its items */

#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		if (first + k < n) {
			out[first + k] = (int32_t) at;
		}

		at += v[k];
	}

	if (blockIdx.x == 0 && threadIdx.x == 0) {
		out[n] = (int32_t) tile_sum[n_tiles];
	}
}

/* d_data[0 .. n] := exclusive prefix sums of d_data[0 .. n) (d_data[n] = total); fails when the total leaves int32 */
int scan_exclusive(int32_t* d_data, int64_t n, int64_t* total) {
	int64_t const n_tiles = (n + kScanTile - 1) / kScanTile;
	int64_t* d_tiles = nullptr;
	int rv = -1;

	*total = 0;

	if (n <= 0) {
		return n == 0 ? bfmg_zero(d_data, sizeof *d_data) : -1;
	}

	if (bfmg_alloc((void**) &d_tiles, ((size_t) n_tiles + 1) * sizeof *d_tiles) < 0) {
		return -1;
	}

	if (
		BFMG_LAUNCH(k_scan_tile_sums, (unsigned) n_tiles, kBlock, 0, (int32_t const*) d_data, n, d_tiles) == 0 &&
		BFMG_LAUNCH(k_scan_tiles, 1, kBlock, 0, d_tiles, n_tiles) == 0 &&
		BFMG_LAUNCH(k_scan_apply, (unsigned) n_tiles, kBlock, 0, (int32_t const*) d_data, d_data, n, (int64_t const*) d_tiles, n_tiles) == 0 &&
		bfmg_download(total, d_tiles + n_tiles, sizeof *total) == 0
	) {
		rv = 0;
	}

	bfmg_free(d_tiles);

	if (rv == 0 && *total > INT32_MAX) {
		bfmg_set_error("symbolic phase: %lld entries do not fit 32-bit indices", (long long) *total);
		return -1;
	}

	return rv;
}

/* ---- sorting one node's segment -------------------------------------------------------------------------- */

constexpr int kLocalKeys = 32;

__device__ __forceinline__ void sift_down(uint64_t* v, int root, int n) {
	uint64_t const x = v[root];

	for (;;) {
		int child = 2 * root + 1;

		if (child >= n) {
			break;
		}

		child += child + 1 < n && v[child + 1] > v[child];

		if (v[child] <= x) {
			break;
		}

		v[root] = v[child];
		root = child;
	}

	v[root] = x;
}

/* ascending, in place; keys are unique */
__device__ void sort_segment(uint64_t* seg, int n) {
	if (n < 2) {
		return;
	}

	if (n <= kLocalKeys) { /* the usual case: insertion sort in local memory (one coalesced wavefront per access) */
		uint64_t loc[kLocalKeys];

		for (int i = 0; i < n; i++) {
			loc[i] = seg[i];
		}

		for (int i = 1; i < n; i++) {
			uint64_t const cur = loc[i];
			int j = i;

			for (; j > 0 && loc[j - 1] > cur; j--) {
				loc[j] = loc[j - 1];
			}

			loc[j] = cur;
		}

		for (int i = 0; i < n; i++) {
			seg[i] = loc[i];
		}

		return;
	}

	for (int i = n / 2 - 1; i >= 0; i--) { /* nodes of high valence: heap sort in place, O(n log n) whatever the input */
		sift_down(seg, i, n);
	}

	for (int end = n - 1; end > 0; end--) {
		uint64_t const top = seg[0];
		seg[0] = seg[end];
		seg[end] = top;
		sift_down(seg, 0, end);
	}
}

/* ---- sparsity pattern ------------------------------------------------------------------------------------ */

__global__ void __launch_bounds__(kBlock) k_sym_count_nodes(int32_t const* __restrict__ elems, int64_t n_inc, int32_t* __restrict__ count) {
	pdl_sync();

	for (int64_t i = (int64_t) blockIdx.x * kBlock + threadIdx.x; i < n_inc; i += (int64_t) gridDim.x * kBlock) {
		atomicAdd(&count[elems[i]], 1);
	}
}

/* every (element, j) writes its `kind` keys into the segment of node elems[e][j]; inc_off[a] = incidences before a */
template <int KIND>
__global__ void __launch_bounds__(kBlock) k_sym_fill_keys(int32_t const* __restrict__ elems, int64_t n_elems, int32_t const* __restrict__ inc_off, int32_t* __restrict__ cursor, uint64_t* __restrict__ keys) {
	pdl_sync();

	for (int64_t i = (int64_t) blockIdx.x * kBlock + threadIdx.x; i < n_elems * KIND; i += (int64_t) gridDim.x * kBlock) {
		int64_t const e = i / KIND;
		int const j = (int) (i - e * KIND);
		int32_t const a = elems[i];
		int64_t const at = ((int64_t) inc_off[a] + atomicAdd(&cursor[a], 1)) * KIND;

#pragma unroll
		for (int k = 0; k < KIND; k++) {
			keys[at + k] = (uint64_t) (uint32_t) elems[e * KIND + k] << 32 | (uint64_t) e << 4 | (uint64_t) (j << 2 | k);
		}
	}
}

/* one thread per node: sort, count distinct columns (a node in no element keeps a lone diagonal block) */
__global__ void __launch_bounds__(kBlock) k_sym_sort_rows(int32_t nb, int kind, int32_t const* __restrict__ inc_off, uint64_t* __restrict__ keys, int32_t* __restrict__ row_len) {
	pdl_sync();

	for (int32_t a = blockIdx.x * kBlock + threadIdx.x; a < nb; a += gridDim.x * kBlock) {
		int64_t const first = (int64_t) inc_off[a] * kind;
		int const cnt = (inc_off[a + 1] - inc_off[a]) * kind;
		uint64_t* const seg = keys + first;

		sort_segment(seg, cnt);

		int len = 0;

		for (int t = 0; t < cnt; t++) {
			len += t == 0 || seg[t] >> 32 != seg[t - 1] >> 32;
		}

		row_len[a] = cnt != 0 ? len : 1;
	}
}

/* width[s] = 32 * longest row of slice s */
__global__ void __launch_bounds__(kBlock) k_sym_slice_width(int32_t nb, int32_t n_slices, int32_t const* __restrict__ row_len, int32_t* __restrict__ width) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);

	for (int32_t s = blockIdx.x * kWarpsPerBlock + threadIdx.x / kWarp; s < n_slices; s += gridDim.x * kWarpsPerBlock) {
		int32_t const a = s * kWarp + lane;
		int longest = a < nb ? row_len[a] : 0;

#pragma unroll
		for (int off = kWarp / 2; off > 0; off >>= 1) {
			longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, off));
		}

		if (lane == 0) {
			width[s] = longest * kWarp;
		}
	}
}

/* every slot starts as padding: it points at its own row (clamped) and has no contributor */
__global__ void __launch_bounds__(kBlock) k_sym_init_slots(int32_t nb, int32_t n_slices, int32_t const* __restrict__ slice_off, int32_t* __restrict__ scol, int32_t* __restrict__ ctr_cnt) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);

	for (int32_t s = blockIdx.x * kWarpsPerBlock + threadIdx.x / kWarp; s < n_slices; s += gridDim.x * kWarpsPerBlock) {
		int32_t const a = min(s * kWarp + lane, nb - 1);

		for (int32_t slot = slice_off[s] + lane; slot < slice_off[s + 1]; slot += kWarp) {
			scol[slot] = a;
			ctr_cnt[slot] = 0;
		}
	}
}

/* one thread per node: the columns of its row, its diagonal slot, contributions per slot */
__global__ void __launch_bounds__(kBlock) k_sym_row_columns(int32_t nb, int kind, int32_t const* __restrict__ inc_off, uint64_t const* __restrict__ keys, int32_t const* __restrict__ slice_off, int32_t* __restrict__ scol, int32_t* __restrict__ diag_pos, int32_t* __restrict__ ctr_cnt) {
	pdl_sync();

	for (int32_t a = blockIdx.x * kBlock + threadIdx.x; a < nb; a += gridDim.x * kBlock) {
		uint64_t const* const seg = keys + (int64_t) inc_off[a] * kind;
		int const cnt = (inc_off[a + 1] - inc_off[a]) * kind;
		int32_t const base = slice_off[a / kWarp] + a % kWarp;
		int32_t diag = base;
		int32_t slot = base - kWarp;
		int32_t run = 0;

		for (int i = 0; i < cnt; i++) {
			uint32_t const b = (uint32_t) (seg[i] >> 32);

			if (i == 0 || b != (uint32_t) (seg[i - 1] >> 32)) {
				if (run != 0) {
					ctr_cnt[slot] = run;
				}

				slot += kWarp;
				run = 0;
				scol[slot] = (int32_t) b;
				diag = b == (uint32_t) a ? slot : diag;
			}

			run++;
		}

		if (run != 0) {
			ctr_cnt[slot] = run;
		}

		diag_pos[a] = diag;
	}
}

/* one thread per node: the packed (element, j, k) lists, slot by slot */
__global__ void __launch_bounds__(kBlock) k_sym_row_contributors(int32_t nb, int kind, int32_t const* __restrict__ inc_off, uint64_t const* __restrict__ keys, int32_t const* __restrict__ slice_off, int32_t const* __restrict__ ctr_ptr, uint32_t* __restrict__ ctr) {
	pdl_sync();

	for (int32_t a = blockIdx.x * kBlock + threadIdx.x; a < nb; a += gridDim.x * kBlock) {
		uint64_t const* const seg = keys + (int64_t) inc_off[a] * kind;
		int const cnt = (inc_off[a + 1] - inc_off[a]) * kind;
		int32_t slot = slice_off[a / kWarp] + a % kWarp - kWarp;
		int32_t fill = 0;

		for (int i = 0; i < cnt; i++) {
			if (i == 0 || seg[i] >> 32 != seg[i - 1] >> 32) {
				slot += kWarp;
				fill = ctr_ptr[slot];
			}

			ctr[fill++] = (uint32_t) seg[i];
		}
	}
}

/* ---- edges ----------------------------------------------------------------------------------------------- */

struct HalfEdge {
	int32_t from, to;
};

__device__ __forceinline__ HalfEdge half_edge(int32_t const* __restrict__ elems, int kind, int64_t h) {
	int64_t const e = h / kind;
	int const j = (int) (h - e * kind);

	return {elems[h], elems[e * kind + (j + 1 == kind ? 0 : j + 1)]};
}

/* segment of a half-edge: its smaller node, counted from the top (the reference's order is descending) */
__global__ void __launch_bounds__(kBlock) k_sym_count_half_edges(int32_t const* __restrict__ elems, int kind, int64_t n_half, int32_t nn, int32_t* __restrict__ count) {
	pdl_sync();

	for (int64_t h = (int64_t) blockIdx.x * kBlock + threadIdx.x; h < n_half; h += (int64_t) gridDim.x * kBlock) {
		HalfEdge const he = half_edge(elems, kind, h);
		atomicAdd(&count[nn - 1 - min(he.from, he.to)], 1);
	}
}

__global__ void __launch_bounds__(kBlock) k_sym_fill_half_edges(int32_t const* __restrict__ elems, int kind, int64_t n_half, int32_t nn, int32_t const* __restrict__ seg_off, int32_t* __restrict__ cursor, uint64_t* __restrict__ keys) {
	pdl_sync();

	for (int64_t h = (int64_t) blockIdx.x * kBlock + threadIdx.x; h < n_half; h += (int64_t) gridDim.x * kBlock) {
		HalfEdge const he = half_edge(elems, kind, h);
		int32_t const seg = nn - 1 - min(he.from, he.to);

		keys[(int64_t) seg_off[seg] + atomicAdd(&cursor[seg], 1)] = (uint64_t) (uint32_t) max(he.from, he.to) << 32 | (uint64_t) h;
	}
}

/* the reference's greedy fusion over one sorted segment.  EMIT = false: sort first, count the edges that come out;
 * EMIT = true: write them (4 x int64: nodes[2], elems[2]) from out_off[segment] on */
template <bool EMIT>
__global__ void __launch_bounds__(kBlock) k_sym_walk_half_edges(int32_t const* __restrict__ elems, int kind, int64_t n_half, int32_t nn, int32_t const* __restrict__ seg_off, uint64_t* __restrict__ keys, int32_t* __restrict__ out_cnt, int32_t const* __restrict__ out_off, int64_t* __restrict__ edges) {
	pdl_sync();

	for (int32_t s = blockIdx.x * kBlock + threadIdx.x; s < nn; s += gridDim.x * kBlock) {
		int32_t const first = seg_off[s];
		int const cnt = seg_off[s + 1] - first;
		uint64_t* const seg = keys + first;

		if (!EMIT) {
			sort_segment(seg, cnt);
		}

		bool const last_segment = (int64_t) first + cnt == n_half; /* later segments are empty */
		int64_t at = EMIT ? out_off[s] : 0;
		int n_out = 0;

		for (int i = 0; i < cnt; i++) {
			int64_t const h = (int64_t) (uint32_t) seg[i];
			HalfEdge const a = half_edge(elems, kind, h);
			int64_t partner = -1;

			if (i + 1 < cnt) {
				int64_t const h2 = (int64_t) (uint32_t) seg[i + 1];
				HalfEdge const b = half_edge(elems, kind, h2);

				if (a.from == b.to && a.to == b.from) {
					partner = h2 / kind;
					i++; /* the partner is consumed */
				}
			}

			else if (last_segment) {
				break; /* the last half-edge of the whole list is only ever emitted as a partner */
			}

			if (EMIT) {
				edges[4 * at + 0] = a.from;
				edges[4 * at + 1] = a.to;
				edges[4 * at + 2] = h / kind;
				edges[4 * at + 3] = partner;
				at++;
			}

			n_out++;
		}

		if (!EMIT) {
			out_cnt[s] = n_out;
		}
	}
}

/* ---- back from the internal numbering -------------------------------------------------------------------- */

__global__ void __launch_bounds__(kBlock) k_gather_blocks(double2* __restrict__ dst, double2 const* __restrict__ src, int32_t const* __restrict__ index, int64_t n) {
	pdl_sync();

	for (int64_t i = (int64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (int64_t) gridDim.x * kBlock) {
		dst[i] = src[index[i]];
	}
}

int grid_for(int64_t items) {
	return bfmg_grid((items + kBlock - 1) / kBlock, 8);
}

} // namespace

/* ---- C ABI ------------------------------------------------------------------------------------------------ */

extern "C" int bfmg_plan_build(int32_t n_nodes, int64_t n_elems, int32_t kind, int32_t const* d_elems, bfmg_pattern_t* pat, int64_t* n_ctr) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (n_nodes <= 0 || (kind != 3 && kind != 4) || n_elems < 0 || n_elems * kind * kind > INT32_MAX) {
		bfmg_set_error("symbolic phase: mesh too large for 32-bit indices");
		return -1;
	}

	int32_t const nb = n_nodes;
	int32_t const n_slices = (nb + kWarp - 1) / kWarp;
	int64_t const n_inc = n_elems * kind;
	int64_t const n_keys = n_inc * kind;

	int32_t *inc_off = nullptr, *cursor = nullptr;
	uint64_t* keys = nullptr;
	int64_t total = 0;
	int rv = -1;

	*pat = bfmg_pattern_t{};
	*n_ctr = 0;

	pat->nb = nb;
	pat->n_slices = n_slices;
	pat->kind = kind;
	pat->row_lo = 0;
	pat->row_hi = nb;

	if (
		bfmg_alloc((void**) &inc_off, ((size_t) nb + 1) * sizeof *inc_off) < 0 || bfmg_alloc((void**) &cursor, ((size_t) nb + 1) * sizeof *cursor) < 0 ||
		bfmg_alloc((void**) &keys, ((size_t) n_keys + 1) * sizeof *keys) < 0 ||
		bfmg_alloc((void**) &pat->row_len, ((size_t) nb + 1) * sizeof(int32_t)) < 0 || bfmg_alloc((void**) &pat->diag_pos, ((size_t) nb + 1) * sizeof(int32_t)) < 0 ||
		bfmg_alloc((void**) &pat->slice_off, ((size_t) n_slices + 1) * sizeof(int32_t)) < 0 ||
		bfmg_zero(inc_off, ((size_t) nb + 1) * sizeof *inc_off) < 0 || bfmg_zero(cursor, ((size_t) nb + 1) * sizeof *cursor) < 0
	) {
		goto done;
	}

	/* incidences per node -> segments of keys, sorted; row lengths */

	if (n_inc > 0 && BFMG_LAUNCH(k_sym_count_nodes, grid_for(n_inc), kBlock, 0, d_elems, n_inc, inc_off) < 0) {
		goto done;
	}

	if (scan_exclusive(inc_off, nb, &total) < 0 || total != n_inc) {
		goto done;
	}

	if (n_inc > 0) {
		int const rc = kind == 3 ?
			BFMG_LAUNCH(k_sym_fill_keys<3>, grid_for(n_inc), kBlock, 0, d_elems, n_elems, (int32_t const*) inc_off, cursor, keys) :
			BFMG_LAUNCH(k_sym_fill_keys<4>, grid_for(n_inc), kBlock, 0, d_elems, n_elems, (int32_t const*) inc_off, cursor, keys);

		if (rc < 0) {
			goto done;
		}
	}

	if (BFMG_LAUNCH(k_sym_sort_rows, grid_for(nb), kBlock, 0, nb, (int) kind, (int32_t const*) inc_off, keys, pat->row_len) < 0) {
		goto done;
	}

	/* slices */

	if (
		BFMG_LAUNCH(k_sym_slice_width, grid_for((int64_t) n_slices * kWarp), kBlock, 0, nb, n_slices, (int32_t const*) pat->row_len, pat->slice_off) < 0 ||
		scan_exclusive(pat->slice_off, n_slices, &total) < 0
	) {
		goto done;
	}

	pat->n_slots = total;

	if (
		bfmg_alloc((void**) &pat->scol, ((size_t) total + 1) * sizeof(int32_t)) < 0 || bfmg_alloc((void**) &pat->ctr_ptr, ((size_t) total + 2) * sizeof(int32_t)) < 0 ||
		bfmg_alloc((void**) &pat->ctr, ((size_t) n_keys + 1) * sizeof(uint32_t)) < 0
	) {
		goto done;
	}

	/* columns, diagonal slots, contributor lists */

	if (
		BFMG_LAUNCH(k_sym_init_slots, grid_for((int64_t) n_slices * kWarp), kBlock, 0, nb, n_slices, (int32_t const*) pat->slice_off, pat->scol, pat->ctr_ptr) < 0 ||
		BFMG_LAUNCH(k_sym_row_columns, grid_for(nb), kBlock, 0, nb, (int) kind, (int32_t const*) inc_off, (uint64_t const*) keys, (int32_t const*) pat->slice_off, pat->scol, pat->diag_pos, pat->ctr_ptr) < 0 ||
		scan_exclusive(pat->ctr_ptr, pat->n_slots, &total) < 0 || total != n_keys ||
		BFMG_LAUNCH(k_sym_row_contributors, grid_for(nb), kBlock, 0, nb, (int) kind, (int32_t const*) inc_off, (uint64_t const*) keys, (int32_t const*) pat->slice_off, (int32_t const*) pat->ctr_ptr, pat->ctr) < 0 ||
		bfmg_sync() < 0
	) {
		goto done;
	}

	*n_ctr = n_keys;
	rv = 0;

done:

	bfmg_free(inc_off);
	bfmg_free(cursor);
	bfmg_free(keys);

	if (rv < 0) {
		bfmg_free(pat->row_len);
		bfmg_free(pat->diag_pos);
		bfmg_free(pat->slice_off);
		bfmg_free(pat->scol);
		bfmg_free(pat->ctr_ptr);
		bfmg_free(pat->ctr);

		*pat = bfmg_pattern_t{};
	}

	return rv;
}

extern "C" int bfmg_edges_build(int32_t n_nodes, int64_t n_elems, int32_t kind, int32_t const* d_elems, int64_t** d_edges, int64_t* n_edges) {
	*d_edges = nullptr;
	*n_edges = 0;

	if (!bfmg_ready()) {
		return -1;
	}

	int64_t const n_half = n_elems * kind;

	if (n_nodes <= 0 || (kind != 3 && kind != 4) || n_elems < 0 || n_half > INT32_MAX) {
		bfmg_set_error("edge derivation: mesh too large for 32-bit indices");
		return -1;
	}

	if (n_half < 2) {
		return 0; /* the reference's loop emits nothing */
	}

	int32_t const nn = n_nodes;
	int32_t *seg_off = nullptr, *cursor = nullptr, *out_off = nullptr;
	uint64_t* keys = nullptr;
	int64_t* edges = nullptr;
	int64_t total = 0;
	int rv = -1;

	if (
		bfmg_alloc((void**) &seg_off, ((size_t) nn + 1) * sizeof *seg_off) < 0 || bfmg_alloc((void**) &cursor, ((size_t) nn + 1) * sizeof *cursor) < 0 ||
		bfmg_alloc((void**) &out_off, ((size_t) nn + 1) * sizeof *out_off) < 0 || bfmg_alloc((void**) &keys, ((size_t) n_half + 1) * sizeof *keys) < 0 ||
		bfmg_zero(seg_off, ((size_t) nn + 1) * sizeof *seg_off) < 0 || bfmg_zero(cursor, ((size_t) nn + 1) * sizeof *cursor) < 0
	) {
		goto done;
	}

	if (
		BFMG_LAUNCH(k_sym_count_half_edges, grid_for(n_half), kBlock, 0, d_elems, (int) kind, n_half, nn, seg_off) < 0 ||
		scan_exclusive(seg_off, nn, &total) < 0 || total != n_half ||
		BFMG_LAUNCH(k_sym_fill_half_edges, grid_for(n_half), kBlock, 0, d_elems, (int) kind, n_half, nn, (int32_t const*) seg_off, cursor, keys) < 0 ||
		BFMG_LAUNCH(k_sym_walk_half_edges<false>, grid_for(nn), kBlock, 0, d_elems, (int) kind, n_half, nn, (int32_t const*) seg_off, keys, out_off, (int32_t const*) nullptr, (int64_t*) nullptr) < 0 ||
		scan_exclusive(out_off, nn, &total) < 0
	) {
		goto done;
	}

	if (total > 0) {
		if (
			bfmg_alloc((void**) &edges, (size_t) total * 4 * sizeof *edges) < 0 ||
			BFMG_LAUNCH(k_sym_walk_half_edges<true>, grid_for(nn), kBlock, 0, d_elems, (int) kind, n_half, nn, (int32_t const*) seg_off, keys, (int32_t*) nullptr, (int32_t const*) out_off, edges) < 0 ||
			bfmg_sync() < 0
		) {
			bfmg_free(edges);
			goto done;
		}
	}

	*d_edges = edges;
	*n_edges = total;
	rv = 0;

done:

	bfmg_free(seg_off);
	bfmg_free(cursor);
	bfmg_free(out_off);
	bfmg_free(keys);

	return rv;
}

extern "C" int bfmg_gather_blocks(double* d_dst, double const* d_src, int32_t const* d_index, size_t n) {
	if (!bfmg_ready()) {
		return -1;
	}

	return n == 0 ? 0 : BFMG_LAUNCH(k_gather_blocks, grid_for((int64_t) n), kBlock, 0, (double2*) d_dst, (double2 const*) d_src, d_index, (int64_t) n);
}
