/*
 * coarse.cuh - device side of the solver's coarse level (included by solver.cu, after Scalars / grid_sum).
 *
 * Two-level additive preconditioner in the Jacobi-scaled variables:  z = r + W E^-1 W^T r,  E = W^T A^ W,
 * W_a = S_a R_a for node a of aggregate g:  S_a = diag(sqrt(a_ii)) (= 1 / dscale),  R_a = [1 0 -dy; 0 1 dx]
 * the rigid-body modes about g's centroid (coarse.c explains the why).  Per PCG iteration:
 *
 *   k_restrict        g = W^T r      one CTA per aggregate over its owned nodes, fixed-order sums
 *   k_coarse_apply    mu = E^-1 g    dense FP64 mat-vec with the explicit inverse, one warp per row;
 *                                    last CTA: rz = r.r + g.mu, beta, rho
 *   k_update_p_coarse p = r + W mu + beta p
 *
 * Setup (once per solve): E column by column through distance-2 colour probing - one plain SpMV per
 * (colour, mode) - then an in-place blocked Gauss-Jordan inversion (E is SPD: no pivoting).  Everything
 * is deterministic: no atomics, fixed reduction trees.  Multi-GPU: restriction partials are all-gathered
 * and folded in rank order, E and E^-1 are replicated (bit-identical on every rank).
 */
#pragma once

namespace {

constexpr int kGjBlock = 32;

struct CoarseWork {
	bfmg_coarse_t C;      /* device pointers of the plan */
	float4* wrow;         /* [nb] (sqrt a_00, sqrt a_11, dx, dy) of every local row, in FP32: W only shapes the
	                       * preconditioner (E is built from the SAME rounded W, so M^-1 stays symmetric positive
	                       * definite); halving its bytes matters, its last digits do not.  Arithmetic stays FP64. */
	double* gpart;        /* [nc] this rank's W^T r */
	double* ggath;        /* [world * nc] all ranks' (several GPUs only) */
	double* g;            /* [nc] the complete W^T r (aliases gpart on one GPU) */
	double* mu;           /* [nc] */
	double* E;            /* [nc * nc] row-major; holds E, then E^-1 */
	double* P;            /* [32 * 32] inverse of the current diagonal block during the inversion */
	int32_t* bad;         /* set when a pivot is not positive: E not SPD, coarse level unusable */
};

__global__ void k_wrow(int nb, double2 const* __restrict__ dscale, double2 const* __restrict__ wgeom, float4* __restrict__ wrow) {
	pdl_sync();

	int const i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < nb) {
		double2 const s = dscale[i];
		double2 const d = wgeom[i];
		wrow[i] = make_float4((float) (1.0 / s.x), (float) (1.0 / s.y), (float) d.x, (float) d.y);
	}
}

/* gpart[3 g + k] = sum over the owned nodes a of aggregate g of (R_a^T S_a v_a)_k; one CTA per aggregate */
__global__ void __launch_bounds__(kBlock) k_restrict(bfmg_coarse_t C, float4 const* __restrict__ wrow, double2 const* __restrict__ v, double* __restrict__ gpart, Scalars* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	__shared__ double part[kWarpsPerBlock][3];

	int const g = blockIdx.x;
	int const beg = C.agg_ptr[g];
	int const end = C.agg_ptr[g + 1];

	double s0 = 0, s1 = 0, s2 = 0;

	for (int i = beg + threadIdx.x; i < end; i += kBlock) {
		int const a = C.agg_nodes[i];
		double2 const vv = v[a];
		float4 const w = wrow[a];

		double const t0 = (double) w.x * vv.x;
		double const t1 = (double) w.y * vv.y;

		s0 += t0;
		s1 += t1;
		s2 += (double) w.z * t1 - (double) w.w * t0;
	}

	s0 = warp_sum(s0);
	s1 = warp_sum(s1);
	s2 = warp_sum(s2);

	if ((threadIdx.x & (kWarp - 1)) == 0) {
		part[threadIdx.x / kWarp][0] = s0;
		part[threadIdx.x / kWarp][1] = s1;
		part[threadIdx.x / kWarp][2] = s2;
	}

	__syncthreads();

	bool const p2p = S->world > 1 && S->use_p2p;
	P2p const& X = S->X;
	uint64_t const round = S->ep_coarse + 1;

	if (threadIdx.x < 3) {
		double t = 0;

#pragma unroll
		for (int w = 0; w < kWarpsPerBlock; w++) {
			t += part[w][threadIdx.x];
		}

		gpart[3 * g + threadIdx.x] = t;

		if (p2p) { /* straight into slot [me] of every peer's coarse channel */
			for (int r = 0; r < X.world; r++) {
				p2p_store_f64((double*) (X.box[r] + X.L.coarse_val) + ((round & 1) * X.world + X.me) * (size_t) X.L.coarse_cap + 3 * g + threadIdx.x, t);
			}

			__threadfence_system();
		}
	}

	/* several GPUs: this rank's share of the pending scalar reduction (r.r) rides along in slot nc, so that
	 * one exchange serves both */

	if (g == 0 && threadIdx.x == 32 && S->world > 1) {
		gpart[C.nc] = S->part;

		if (p2p) {
			for (int r = 0; r < X.world; r++) {
				p2p_store_f64((double*) (X.box[r] + X.L.coarse_val) + ((round & 1) * X.world + X.me) * (size_t) X.L.coarse_cap + C.nc, S->part);
			}

			__threadfence_system();
		}
	}

	if (!p2p) {
		return;
	}

	/* the last CTA to get here publishes the round */

	__shared__ bool last;

	__syncthreads();

	if (threadIdx.x == 0) {
		__threadfence_system();
		last = atomicAdd(&S->ticket2, 1u) == gridDim.x - 1;
	}

	__syncthreads();

	if (last && threadIdx.x == 0) {
		__threadfence_system();

		for (int r = 0; r < X.world; r++) {
			p2p_publish(X, r, X.L.coarse_seq, round);
		}

		S->ep_coarse = round;
		S->ticket2 = 0;
	}
}

/* several GPUs: g[j] = sum over ranks (in rank order) of ggath[r][j]; rows are kCoarseStride = nc + 8 long,
 * slot nc holding the ranks' shares of r.r, folded here too when FOLD_RR (k_fold<kFoldRr>'s job otherwise) */
template <bool FOLD_RR>
__global__ void k_coarse_fold(int nc, int n_real, int world, double const* __restrict__ ggath, double* __restrict__ g, Scalars* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	int const j = blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t) nc + 8;

	if (S->use_p2p) { /* the ranks' posts are in MY mailbox: wait for the round, then read them from there */
		P2p const& X = S->X;
		uint64_t const round = S->ep_coarse;

		if (threadIdx.x < world) {
			p2p_wait(X, p2p_seq(X, X.me, X.L.coarse_seq, round, threadIdx.x), round);
		}

		__syncthreads();

		ggath = (double const*) (X.box[X.me] + X.L.coarse_val) + (round & 1) * (size_t) X.world * X.L.coarse_cap;
		stride = (size_t) X.L.coarse_cap;
	}

	if (j < nc) {
		double t = 0;

		/* only the 3 n_agg real entries are ever posted; the padding up to nc is zero by definition (a mailbox
		 * keeps what an earlier, larger solve left there) */

		for (int r = 0; r < world && j < n_real; r++) {
			t += __ldcg(&ggath[r * stride + j]);
		}

		g[j] = t;
	}

	if (FOLD_RR && j == nc && !S->done) {
		double total = 0;

		for (int r = 0; r < world; r++) {
			total += __ldcg(&ggath[r * stride + nc]);
		}

		fold<kFoldRr>(S, total);
	}
}

/* mu = E^-1 g, then in the last CTA: rz = r.r + g.mu;  beta = rz / rho (0 when FIRST);  rho = rz.
 * A CTA takes kCoarseRows rows, four warps per row (a quarter of the row each, four independent FMA
 * chains per lane): E^-1 is 75 MB at nc = 3072 and one chain per row would be pure load latency.
 * Partial sums are combined in a fixed order. */
constexpr int kCoarseRows = kWarpsPerBlock / 4;

/* DIST (several GPUs with peer memory): this rank applies rows [row0, row0 + gridDim.x * kCoarseRows) only and
 * stores its part of mu into every peer's mailbox; k_coarse_finish then does the scalar part */
template <bool FIRST, bool DIST>
__global__ void __launch_bounds__(kBlock) k_coarse_apply(int nc, int row0, double const* __restrict__ Einv, double const* __restrict__ g, double* __restrict__ mu, double* __restrict__ partials, Scalars* S) {
	pdl_sync();

	if (!FIRST && S->done) {
		return;
	}

	__shared__ double quarter_sum[kWarpsPerBlock];

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = threadIdx.x / kWarp;
	int const row = row0 + blockIdx.x * kCoarseRows + warp / 4;
	int const span = ((nc + 3) / 4 + kWarp - 1) / kWarp * kWarp; /* columns per quarter, a multiple of 32 */
	int const j_end = min(nc, (warp % 4 + 1) * span);

	double t = 0;

	if (row < nc) {
		double const* const e = Einv + (size_t) row * nc;
		double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
		int j = (warp % 4) * span + lane;

		for (; j + 3 * kWarp < j_end; j += 4 * kWarp) {
			t0 = fma(ld_stream(&e[j]), __ldg(&g[j]), t0);
			t1 = fma(ld_stream(&e[j + kWarp]), __ldg(&g[j + kWarp]), t1);
			t2 = fma(ld_stream(&e[j + 2 * kWarp]), __ldg(&g[j + 2 * kWarp]), t2);
			t3 = fma(ld_stream(&e[j + 3 * kWarp]), __ldg(&g[j + 3 * kWarp]), t3);
		}

		for (; j < j_end; j += kWarp) {
			t0 = fma(ld_stream(&e[j]), __ldg(&g[j]), t0);
		}

		t = warp_sum((t0 + t1) + (t2 + t3));
	}

	if (lane == 0) {
		quarter_sum[warp] = t;
	}

	__syncthreads();

	double acc = 0;

	if (lane == 0 && warp % 4 == 0 && row < nc) {
		double const m = (quarter_sum[warp] + quarter_sum[warp + 1]) + (quarter_sum[warp + 2] + quarter_sum[warp + 3]);

		if (DIST) {
			P2p const& X = S->X;
			uint64_t const round = S->ep_mu + 1;

			for (int r = 0; r < X.world; r++) {
				p2p_store_f64((double*) (X.box[r] + X.L.mu_val) + (round & 1) * (size_t) X.L.coarse_cap + row, m);
			}

			__threadfence_system();
		}

		else {
			mu[row] = m;
			acc = m * g[row];
		}
	}

	double total;

	if (grid_sum(acc, partials, &S->ticket, &total) && threadIdx.x == 0) {
		if (DIST) { /* every CTA of this rank has stored and fenced its rows: publish the round */
			P2p const& X = S->X;
			uint64_t const round = S->ep_mu + 1;

			__threadfence_system();

			for (int r = 0; r < X.world; r++) {
				p2p_publish(X, r, X.L.mu_seq, round);
			}

			S->ep_mu = round;
		}

		else {
			double const rz = S->rr + total;

			S->beta = FIRST ? 0 : rz / S->rho;
			S->rho = rz;
		}
	}
}

/* DIST: all parts of mu are in MY mailbox once every rank's round is there; copy them out (mu) and finish
 * the scalars: rz = r.r + g.mu, beta, rho.  One CTA. */
template <bool FIRST>
__global__ void __launch_bounds__(kBlock) k_coarse_finish(int nc, double const* __restrict__ g, double* __restrict__ mu, Scalars* S) {
	pdl_sync();

	if (!FIRST && S->done) {
		return;
	}

	__shared__ double warp_part[kWarpsPerBlock];

	P2p const& X = S->X;
	uint64_t const round = S->ep_mu;

	if (threadIdx.x < X.world) {
		p2p_wait(X, p2p_seq(X, X.me, X.L.mu_seq, round, threadIdx.x), round);
	}

	__syncthreads();

	double const* const src = (double const*) (X.box[X.me] + X.L.mu_val) + (round & 1) * (size_t) X.L.coarse_cap;
	double acc = 0;

	for (int i = threadIdx.x; i < nc; i += kBlock) {
		double const m = __ldcg(&src[i]);

		mu[i] = m;
		acc = fma(m, g[i], acc);
	}

	acc = warp_sum(acc);

	if ((threadIdx.x & (kWarp - 1)) == 0) {
		warp_part[threadIdx.x / kWarp] = acc;
	}

	__syncthreads();

	if (threadIdx.x == 0) {
		double total = 0;

#pragma unroll
		for (int w = 0; w < kWarpsPerBlock; w++) {
			total += warp_part[w];
		}

		double const rz = S->rr + total;

		S->beta = FIRST ? 0 : rz / S->rho;
		S->rho = rz;
	}
}

/* p = r + W mu + beta p */
__global__ void __launch_bounds__(kBlock) k_update_p_coarse(int n2, int row0, bfmg_coarse_t C, float4 const* __restrict__ wrow, double const* __restrict__ mu, double2 const* __restrict__ r, double2* __restrict__ p, Scalars const* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	double const beta = S->beta;

	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
		int const a = row0 + i;
		int const g = C.agg[a];
		float4 const w = wrow[a];
		double2 const rv = r[a];
		double2 pv = p[a];

		double const m0 = __ldg(&mu[3 * g + 0]);
		double const m1 = __ldg(&mu[3 * g + 1]);
		double const m2 = __ldg(&mu[3 * g + 2]);

		double const z0 = rv.x + (double) w.x * (m0 - (double) w.w * m2);
		double const z1 = rv.y + (double) w.y * (m1 + (double) w.z * m2);

		pv.x = fma(beta, pv.x, z0);
		pv.y = fma(beta, pv.y, z1);

		p[a] = pv;
	}
}

/* ---- setup: probing ------------------------------------------------------------------------------ */

/* v = mode m of every aggregate of colour c, 0 elsewhere (all local rows, ghosts included: no exchange needed) */
__global__ void k_probe_vector(int nb, bfmg_coarse_t C, float4 const* __restrict__ wrow, int color, int mode, double2* __restrict__ v) {
	pdl_sync();

	int const a = blockIdx.x * blockDim.x + threadIdx.x;

	if (a >= nb) {
		return;
	}

	double2 out = make_double2(0, 0);

	if (C.color[C.agg[a]] == color) {
		float4 const w = wrow[a];

		out.x = mode == 0 ? (double) w.x : (mode == 2 ? -((double) w.x * (double) w.w) : 0);
		out.y = mode == 1 ? (double) w.y : (mode == 2 ? (double) w.y * (double) w.z : 0);
	}

	v[a] = out;
}

/* g = W^T A^ (modes m of colour c): rows 3h..3h+2 are column 3 src + m of E, src = the colour-c aggregate near h */
__global__ void k_probe_scatter(bfmg_coarse_t C, int color, int mode, double const* __restrict__ g, double* __restrict__ E) {
	pdl_sync();

	int const h = blockIdx.x * blockDim.x + threadIdx.x;

	if (h >= C.n_agg) {
		return;
	}

	int const src = C.color_nbr[(size_t) h * C.n_colors + color];

	if (src < 0) {
		return;
	}

	for (int k = 0; k < 3; k++) {
		E[(size_t) (3 * h + k) * C.nc + 3 * src + mode] = g[3 * h + k];
	}
}

/* identity on the padding rows (3 n_agg .. nc) */
__global__ void k_coarse_pad(bfmg_coarse_t C, double* __restrict__ E) {
	pdl_sync();

	int const i = 3 * C.n_agg + blockIdx.x * blockDim.x + threadIdx.x;

	if (i < C.nc) {
		E[(size_t) i * C.nc + i] = 1;
	}
}

/* ---- setup: in-place blocked Gauss-Jordan inversion of the SPD matrix E (n x n, n % 32 == 0) ----------
 *
 * For block k (rows/columns K):  P = E_KK^-1;  E_K* = P E_K*;  E_ij -= E_iK E_Kj (i, j not in K);
 * E_*K = -E_*K P;  E_KK = P.  After the last block E holds its inverse. */

__global__ void __launch_bounds__(1024) k_gj_diag(int n, int k, double const* __restrict__ E, double* __restrict__ P, int32_t* bad) {
	pdl_sync();

	__shared__ double a[kGjBlock][kGjBlock + 1];

	int const i = threadIdx.x / kGjBlock;
	int const j = threadIdx.x % kGjBlock;
	int const base = k * kGjBlock;

	a[i][j] = E[(size_t) (base + i) * n + base + j];
	__syncthreads();

	for (int t = 0; t < kGjBlock; t++) {
		double const piv = a[t][t];

		if (!(piv > 0) && threadIdx.x == 0) {
			*bad = 1;
		}

		double const inv = 1.0 / piv;
		double const ait = a[i][t];
		double const atj = a[t][j];

		__syncthreads();

		double v;

		if (i == t) {
			v = j == t ? inv : atj * inv;
		}

		else {
			v = j == t ? -ait * inv : a[i][j] - ait * (atj * inv);
		}

		a[i][j] = v;
		__syncthreads();
	}

	P[i * kGjBlock + j] = a[i][j];
}

/* row panel: E[K, J] = P * E[K, J] for every column block J != k; one CTA per J */
__global__ void __launch_bounds__(1024) k_gj_row(int n, int k, int lim, double* __restrict__ E, double const* __restrict__ P) {
	pdl_sync();

	int const J = blockIdx.x;

	if (J == k || J * kGjBlock >= lim) {
		return;
	}

	__shared__ double p[kGjBlock][kGjBlock + 1];
	__shared__ double a[kGjBlock][kGjBlock + 1];

	int const i = threadIdx.x / kGjBlock;
	int const j = threadIdx.x % kGjBlock;

	double* const tile = E + (size_t) (k * kGjBlock) * n + J * kGjBlock;

	p[i][j] = P[i * kGjBlock + j];
	a[i][j] = tile[(size_t) i * n + j];
	__syncthreads();

	double s = 0;

#pragma unroll
	for (int t = 0; t < kGjBlock; t++) {
		s = fma(p[i][t], a[t][j], s);
	}

	tile[(size_t) i * n + j] = s;
}

/* where the pivot row panel R = P E[K, :] (32 x n) and P (32 x 32) of a Gauss-Jordan step come from: the matrix
 * itself on one GPU; this rank's mailbox when the inversion is distributed by row blocks (p2p.cuh, panel channel) */
struct GjPanel {
	bool dist;
	int owner;        /* rank that owns pivot block k */
	uint64_t round;   /* k + 1 */
};

__device__ __forceinline__ double const* gj_panel(Scalars const* S, GjPanel const& G) {
	P2p const& X = S->X;
	return (double const*) (X.box[X.me] + X.L.panel_val) + (G.round & 1) * (32 * (size_t) X.L.coarse_cap + 32 * 32 + 8);
}

/* trailing update: E[i, j] -= sum_t E[i, K_t] * R[t, j] for the rows i of [row_lo, row_hi) and all j, both outside K;
 * 64 x 64 tile per CTA, 4 x 4 per thread.
 * (A 128 x 128 / 8 x 8-per-thread variant was measured SLOWER - 168 registers, one CTA per SM: 620 us against 380 us
 * per launch at n = 6304 - so the small tile stays.) */
__global__ void __launch_bounds__(256, 4) k_gj_update(int n, int k, int lim, double* __restrict__ E, int row_lo, int row_hi, GjPanel G, Scalars* S) {
	pdl_sync();

	__shared__ double col[64][kGjBlock + 1]; /* E[I, K] */
	__shared__ double row[kGjBlock][64 + 1]; /* R[K, J] */

	int const i0 = row_lo + blockIdx.y * 64;
	int const j0 = blockIdx.x * 64;
	int const kb = k * kGjBlock;

	if (i0 >= lim || j0 >= lim) {
		return; /* E[i, K] or R[K, j] is still all zero there: E is banded and pivots 0..k have filled rows and
		         * columns below lim = 32 (k + 1) + half_bw only (see coarse_invert) */
	}

	double const* R = E + (size_t) kb * n;

	if (G.dist) { /* the owner's panel lands in my mailbox: wait for this step's round */
		P2p const& X = S->X;

		if (threadIdx.x == 0) {
			p2p_wait(X, p2p_seq(X, X.me, X.L.panel_seq, G.round, G.owner), G.round);
		}

		__syncthreads();
		R = gj_panel(S, G);
	}

	for (int t = threadIdx.x; t < 64 * kGjBlock; t += 256) {
		int const r = t / kGjBlock, c = t % kGjBlock;
		col[r][c] = i0 + r < row_hi ? E[(size_t) (i0 + r) * n + kb + c] : 0;
	}

	for (int t = threadIdx.x; t < kGjBlock * 64; t += 256) {
		int const r = t / 64, c = t % 64;
		row[r][c] = j0 + c < n ? __ldcg(&R[(size_t) r * n + j0 + c]) : 0;
	}

	__syncthreads();

	int const ti = (threadIdx.x / 16) * 4;
	int const tj = (threadIdx.x % 16) * 4;

	double acc[4][4] = {};

#pragma unroll 8
	for (int t = 0; t < kGjBlock; t++) {
		double cv[4], rv[4];

#pragma unroll
		for (int u = 0; u < 4; u++) {
			cv[u] = col[ti + u][t];
			rv[u] = row[t][tj + u];
		}

#pragma unroll
		for (int u = 0; u < 4; u++) {
#pragma unroll
			for (int w = 0; w < 4; w++) {
				acc[u][w] = fma(cv[u], rv[w], acc[u][w]);
			}
		}
	}

#pragma unroll
	for (int u = 0; u < 4; u++) {
		int const i = i0 + ti + u;

		if (i >= row_hi || (i >= kb && i < kb + kGjBlock)) {
			continue;
		}

#pragma unroll
		for (int w = 0; w < 4; w++) {
			int const j = j0 + tj + w;

			if (j < n && !(j >= kb && j < kb + kGjBlock)) {
				E[(size_t) i * n + j] -= acc[u][w];
			}
		}
	}
}

/* column panel: E[I, K] = -E[I, K] * P for the row blocks I != k of this rank;  E[K, K] = P where it owns K.
 * One CTA per row block from row_lo on.  Distributed: P comes from the mailbox, the owner's "not positive
 * definite" verdict is taken over, and the last CTA acknowledges the step to every peer. */
__global__ void __launch_bounds__(1024) k_gj_col(int n, int k, int lim, double* __restrict__ E, double const* __restrict__ Plocal, int row_lo, GjPanel G, int32_t* bad, Scalars* S) {
	pdl_sync();

	int const I = row_lo / kGjBlock + blockIdx.x;

	__shared__ double p[kGjBlock][kGjBlock + 1];
	__shared__ double a[kGjBlock][kGjBlock + 1];

	int const i = threadIdx.x / kGjBlock;
	int const j = threadIdx.x % kGjBlock;

	double const* P = Plocal;

	if (G.dist) {
		/* k_gj_update waited for this round too, but its CTAs beyond lim return before they do */

		if (threadIdx.x == 0) {
			P2p const& X = S->X;
			p2p_wait(X, p2p_seq(X, X.me, X.L.panel_seq, G.round, G.owner), G.round);
		}

		__syncthreads();

		P = gj_panel(S, G) + 32 * (size_t) n;

		if (blockIdx.x == 0 && threadIdx.x == 0 && __ldcg(&P[kGjBlock * kGjBlock]) != 0) {
			*bad = 1;
		}
	}

	double* const tile = E + (size_t) (I * kGjBlock) * n + k * kGjBlock;

	p[i][j] = __ldcg(&P[i * kGjBlock + j]);
	a[i][j] = tile[(size_t) i * n + j];
	__syncthreads();

	if (I == k) {
		tile[(size_t) i * n + j] = p[i][j];
	}

	else if (I * kGjBlock < lim) { /* beyond lim the column panel is zero and stays zero */
		double s = 0;

#pragma unroll
		for (int t = 0; t < kGjBlock; t++) {
			s = fma(a[i][t], p[t][j], s);
		}

		tile[(size_t) i * n + j] = -s;
	}

	if (!G.dist) {
		return;
	}

	/* acknowledge: this rank has finished step k (its panel buffer may be overwritten two steps from now) */

	__shared__ bool last;

	__syncthreads();

	if (threadIdx.x == 0) {
		__threadfence();
		last = atomicAdd(&S->ticket2, 1u) == gridDim.x - 1;
	}

	__syncthreads();

	if (last && threadIdx.x == 0) {
		P2p const& X = S->X;

		for (int r = 0; r < X.world; r++) {
			p2p_store_u64((uint64_t*) (X.box[r] + X.L.gj_done) + X.me, G.round);
		}

		S->ticket2 = 0;
	}
}

/* distributed: the owner of pivot block k stores R = E[K, :] (already multiplied by P), P and its verdict into
 * every rank's panel buffer of this round, then publishes the round */
__global__ void __launch_bounds__(kBlock) k_gj_bcast(int n, int k, double const* __restrict__ E, double const* __restrict__ P, int32_t const* __restrict__ bad, GjPanel G, Scalars* S) {
	pdl_sync();

	P2p const& X = S->X;

	/* the buffer of this parity was last read in step k - 2: every rank must have acknowledged it */

	if (threadIdx.x < X.world && G.round >= 3) {
		p2p_wait(X, (uint64_t const*) (X.box[X.me] + X.L.gj_done) + threadIdx.x, G.round - 2);
	}

	__syncthreads();

	size_t const panel = 32 * (size_t) X.L.coarse_cap + 32 * 32 + 8;
	size_t const count = 32 * (size_t) n + 32 * 32 + 1;
	double const* const src = E + (size_t) k * kGjBlock * n;

	for (size_t t = blockIdx.x * (size_t) blockDim.x + threadIdx.x; t < count; t += (size_t) gridDim.x * blockDim.x) {
		double const v = t < 32 * (size_t) n ? src[t] : (t < 32 * (size_t) n + 32 * 32 ? P[t - 32 * (size_t) n] : (double) *bad);

		for (int r = 0; r < X.world; r++) {
			p2p_store_f64((double*) (X.box[r] + X.L.panel_val) + (G.round & 1) * panel + t, v);
		}
	}

	__threadfence_system();

	__shared__ bool last;

	__syncthreads();

	if (threadIdx.x == 0) {
		__threadfence_system();
		last = atomicAdd(&S->ticket2, 1u) == gridDim.x - 1;
	}

	__syncthreads();

	if (last && threadIdx.x == 0) {
		__threadfence_system();

		for (int r = 0; r < X.world; r++) {
			p2p_publish(X, r, X.L.panel_seq, G.round);
		}

		S->ticket2 = 0;
	}
}

__global__ void k_gj_ack_all(uint64_t rounds, Scalars* S) {
	pdl_sync();

	P2p const& X = S->X;

	if (threadIdx.x < X.world) {
		p2p_store_u64((uint64_t*) (X.box[threadIdx.x] + X.L.gj_done) + X.me, rounds);
	}
}

/* In place.  Replicated (one GPU, or NCCL exchanges): every rank inverts all of E.  Distributed (peer memory):
 * rank r owns the row blocks [r * blocks_per, (r + 1) * blocks_per) - the rows of E^-1 it will apply - and only
 * updates those; the pivot panel of each step comes from its owner through the mailboxes. */
int coarse_invert(CoarseWork const& W, Scalars* S, bool dist, int rank, int blocks_per) {
	int const n = W.C.nc;
	int const blocks = n / kGjBlock;

	int const my_b0 = dist ? (rank * blocks_per < blocks ? rank * blocks_per : blocks) : 0;
	int const my_b1 = dist ? ((rank + 1) * blocks_per < blocks ? (rank + 1) * blocks_per : blocks) : blocks;
	int const row_lo = my_b0 * kGjBlock;
	int const row_hi = my_b1 * kGjBlock;

	dim3 const tiles((n + 63) / 64, (row_hi - row_lo + 63) / 64);

	/* a rank without rows never reads a panel: it acknowledges every step up front */

	if (dist && row_hi == row_lo && BFMG_LAUNCH(k_gj_ack_all, 1, kWarp, 0, (uint64_t) blocks + 2, S) < 0) {
		return -1;
	}

	for (int k = 0; k < blocks; k++) {
		GjPanel G;

		G.dist = dist;
		G.owner = dist ? k / blocks_per : 0;
		G.round = (uint64_t) k + 1;

		bool const mine = !dist || G.owner == rank;

		/* E starts banded (adjacent aggregates have close numbers) and pivot blocks 0..k fill rows and columns
		 * below lim only: everything at or beyond lim is still zero in both panels - a third of the n^3 work */
		int const lim = (k + 1) * kGjBlock + W.C.half_bw < n ? (k + 1) * kGjBlock + W.C.half_bw : n;

		if (mine && (
			BFMG_LAUNCH(k_gj_diag, 1, 1024, 0, n, k, W.E, W.P, W.bad) < 0 ||
			BFMG_LAUNCH(k_gj_row, blocks, 1024, 0, n, k, lim, W.E, W.P) < 0 ||
			(dist && BFMG_LAUNCH(k_gj_bcast, 64, kBlock, 0, n, k, W.E, W.P, W.bad, G, S) < 0)
		)) {
			return -1;
		}

		if (row_hi > row_lo && (
			BFMG_LAUNCH(k_gj_update, tiles, 256, 0, n, k, lim, W.E, row_lo, row_hi, G, S) < 0 ||
			BFMG_LAUNCH(k_gj_col, my_b1 - my_b0, 1024, 0, n, k, lim, W.E, W.P, row_lo, G, W.bad, S) < 0
		)) {
			return -1;
		}
	}

	return 0;
}

} // namespace
