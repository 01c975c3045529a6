/*
 * batch.cu - small systems: one CTA runs the WHOLE FP64 Jacobi-PCG of one system, many systems per launch.
 *
 * BASELINE.json configs[4] ("1024 batched small systems, one CTA group per system") extends the loop of
 * examples/benchmark.py:10-18, which calls sim.run() ten times on the 670-DOF mesh and rebuilds the
 * whole system each time (sim.c:112).  The same kernel is also the path bfm_sim_run takes for ONE small
 * mesh (the reference's own examples are all small: 670, 2 718 and 3 326 DOF): on those the three-kernel
 * iteration of solver.cu is pure launch latency (~15-20 us per iteration), while a CTA that keeps x, r, p,
 * q in shared memory iterates in ~1-2 us.
 *
 * Layout: the systems are consecutive block-row ranges [row_lo, row_hi) of ONE SELL-32 pattern (a batch
 * is assembled as one mesh of disconnected components, each padded to a multiple of 32 nodes so that no
 * slice straddles two systems - job.c).  Per system the CTA does, with the arithmetic of solver.cu:
 *     load b^, r = p = b^, x = 0;  loop { q = A^ p;  alpha;  x, r;  beta;  p }  until ||r|| <= tol ||b^||;
 *     recompute b^ - A^ x^ and the backward error;  write x = D^-1/2 x^ and the per-system status.
 * The Jacobi-scaled matrix A^ and b^ come from k_jacobi / k_scale_matrix (one launch each for the whole
 * batch).  Matrix values and columns are streamed from L1/L2 every iteration (a 670-DOF system is
 * 108 KB, 1024 of them 110 MB - inside the 126 MB L2); the four vectors live in shared memory.
 * Measured dead ends (round 1): keeping the matrix itself in shared memory (no faster for one system, slower for a
 * batch - one CTA per SM), a tree fold of the warp partials and a deeper unroll of the slot loop (both slightly
 * slower): an iteration is bound by its chain of three barriers and the 11-of-32-warp SpMV phase, not by the stream.
 * Reductions are fixed-order (shuffle tree, then warp partials summed by every thread in warp order):
 * deterministic, identical in every thread, so the convergence branch is CTA-uniform.
 */
#include "gpu_internal.cuh"

#include <cmath>

namespace {

template <int THREADS>
__device__ __forceinline__ double cta_sum(double v, double* warp_part) {
	constexpr int kWarps = THREADS / kWarp;

	v = warp_sum(v);

	if ((threadIdx.x & (kWarp - 1)) == 0) {
		warp_part[threadIdx.x / kWarp] = v;
	}

	__syncthreads();

	double s = 0;

#pragma unroll
	for (int w = 0; w < kWarps; w++) {
		s += warp_part[w];
	}

	return s;
}

/* y = A^ v over the block rows [lo, hi) of this CTA's system; v is the CTA's shared-memory vector
 * (indexed from lo); calls f(row - lo, y0, y1) for every real row */
template <int THREADS, typename F>
__device__ __forceinline__ void cta_spmv(bfmg_pattern_t const& P, double2 const* __restrict__ stop, double2 const* __restrict__ sbot, int lo, int hi, double2 const* v, F&& f) {
	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = threadIdx.x / kWarp;

	for (int slice = lo / kWarp + warp; slice < (hi + kWarp - 1) / kWarp; slice += THREADS / kWarp) {
		int const row = slice * kWarp + lane;
		int const beg = __ldg(&P.slice_off[slice]);
		int const end = __ldg(&P.slice_off[slice + 1]);

		double y0 = 0, y1 = 0;

#pragma unroll 2
		for (int slot = beg + lane; slot < end; slot += kWarp) {
			int const col = __ldg(&P.scol[slot]);
			double2 const t = __ldg(&stop[slot]);
			double2 const u = __ldg(&sbot[slot]);
			double2 const xv = v[col - lo];

			y0 = fma(t.x, xv.x, fma(t.y, xv.y, y0));
			y1 = fma(u.x, xv.x, fma(u.y, xv.y, y1));
		}

		if (row < hi) {
			f(row - lo, y0, y1);
		}
	}
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_pcg_cta(
	const __grid_constant__ bfmg_pattern_t P, double2 const* __restrict__ stop, double2 const* __restrict__ sbot,
	double2 const* __restrict__ bhat, double2 const* __restrict__ dscale, double2* __restrict__ x_out,
	bfmg_batch_range_t const* __restrict__ ranges, bfmg_batch_status_t* __restrict__ status, double tol, int max_iter
) {
	pdl_sync();

	extern __shared__ double2 smem[];
	__shared__ double part[2][THREADS / kWarp];

	int const sys = blockIdx.x;
	int const lo = ranges[sys].row_lo;
	int const hi = ranges[sys].row_hi;
	int const n = hi - lo;

	double2* const x = smem;
	double2* const r = x + n;
	double2* const p = r + n;
	double2* const q = p + n;

	double acc = 0;

	for (int i = threadIdx.x; i < n; i += THREADS) {
		double2 const v = bhat[lo + i];

		x[i] = make_double2(0, 0);
		r[i] = v;
		p[i] = v;

		acc = fma(v.x, v.x, fma(v.y, v.y, acc));
	}

	double rho = cta_sum<THREADS>(acc, part[0]);
	double const bnorm2 = rho;
	double const tol2 = tol * tol;

	int iter = 0;
	int done = (rho == 0) ? 1 : (rho == rho ? 0 : 2); /* as solver.cu: 1 converged, 2 breakdown, 3 iteration limit */

	__syncthreads(); /* p visible to every warp; part[0] may be reused after the next barrier */

	while (!done) {
		/* q = A^ p, p.q */

		acc = 0;

		cta_spmv<THREADS>(P, stop, sbot, lo, hi, p, [&](int i, double y0, double y1) {
			double2 const pv = p[i];
			q[i] = make_double2(y0, y1);
			acc = fma(pv.x, y0, fma(pv.y, y1, acc));
		});

		double const pq = cta_sum<THREADS>(acc, part[1]);

		if (pq == 0 || !(pq == pq) || isinf(pq)) {
			done = 2;
			break;
		}

		double const alpha = rho / pq;

		/* x += alpha p, r -= alpha q, r.r  (q was written by other lanes: the barrier inside cta_sum ordered it) */

		acc = 0;

		for (int i = threadIdx.x; i < n; i += THREADS) {
			double2 const pv = p[i];
			double2 const qv = q[i];
			double2 xv = x[i];
			double2 rv = r[i];

			xv.x = fma(alpha, pv.x, xv.x);
			xv.y = fma(alpha, pv.y, xv.y);
			rv.x = fma(-alpha, qv.x, rv.x);
			rv.y = fma(-alpha, qv.y, rv.y);

			x[i] = xv;
			r[i] = rv;

			acc = fma(rv.x, rv.x, fma(rv.y, rv.y, acc));
		}

		double const rr = cta_sum<THREADS>(acc, part[0]);
		double const beta = rr / rho;

		rho = rr;
		iter++;

		if (!(rr == rr)) {
			done = 2;
		}

		else if (rr <= tol2 * bnorm2) {
			done = 1;
		}

		else if (iter >= max_iter) {
			done = 3;
		}

		else { /* p = r + beta p, same row-to-thread map as the update above */
			for (int i = threadIdx.x; i < n; i += THREADS) {
				double2 const rv = r[i];
				double2 pv = p[i];

				pv.x = fma(beta, pv.x, rv.x);
				pv.y = fma(beta, pv.y, rv.y);

				p[i] = pv;
			}
		}

		__syncthreads();
	}

	/* verification: true residual b^ - A^ x^ and ||x^|| (solver.cu explains the backward-error criterion) */

	__syncthreads();

	double res2 = 0, xn2 = 0;

	if (bnorm2 > 0) {
		acc = 0;

		cta_spmv<THREADS>(P, stop, sbot, lo, hi, x, [&](int i, double y0, double y1) {
			double2 const bb = bhat[lo + i];
			double const r0 = bb.x - y0;
			double const r1 = bb.y - y1;
			acc = fma(r0, r0, fma(r1, r1, acc));
		});

		res2 = cta_sum<THREADS>(acc, part[1]);

		acc = 0;

		for (int i = threadIdx.x; i < n; i += THREADS) {
			double2 const xv = x[i];
			acc = fma(xv.x, xv.x, fma(xv.y, xv.y, acc));
		}

		xn2 = cta_sum<THREADS>(acc, part[0]);
	}

	for (int i = threadIdx.x; i < n; i += THREADS) {
		double2 const s = dscale[lo + i];
		double2 const xv = x[i];
		x_out[lo + i] = make_double2(s.x * xv.x, s.y * xv.y);
	}

	if (threadIdx.x == 0) {
		bfmg_batch_status_t st;

		st.iterations = iter;
		st.converged = done == 1 ? 1 : (done == 3 ? 0 : -1);
		st.rel_residual = bnorm2 > 0 ? sqrt(rho / bnorm2) : 0;
		st.true_rel_residual = bnorm2 > 0 ? sqrt(res2 / bnorm2) : 0;
		st.backward_error = bnorm2 > 0 ? sqrt(res2) / (sqrt(xn2) + sqrt(bnorm2)) : 0;

		status[sys] = st;
	}
}

} // namespace

extern "C" {

int bfmg_batch_max_rows(void) {
	/* four double2 vectors in at most 200 KB of the 227 KB a CTA may use */
	return (200 * 1024) / (4 * (int) sizeof(double2));
}

int bfmg_pcg_batch(bfmg_pattern_t const* pat, double const* d_val, double const* d_b, double* d_x, bfmg_pcg_opts_t const* opts, int32_t n_sys, bfmg_batch_range_t const* ranges, bfmg_batch_status_t* status, float* ms) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (n_sys <= 0) {
		return 0;
	}

	int max_rows = 0;

	for (int32_t s = 0; s < n_sys; s++) {
		int const rows = ranges[s].row_hi - ranges[s].row_lo;

		if (rows < 0 || ranges[s].row_lo % kWarp != 0 || ranges[s].row_hi > pat->nb) {
			bfmg_set_error("batch system %d has a bad row range [%d, %d)", s, ranges[s].row_lo, ranges[s].row_hi);
			return -1;
		}

		max_rows = rows > max_rows ? rows : max_rows;
	}

	if (max_rows > bfmg_batch_max_rows()) {
		bfmg_set_error("a system of %d node rows does not fit the one-CTA solver (limit %d)", max_rows, bfmg_batch_max_rows());
		return -1;
	}

	size_t const vec_bytes = (size_t) pat->nb * sizeof(double2);
	size_t const mat_bytes = (size_t) pat->n_slots * 2 * sizeof(double2);
	size_t const range_bytes = (size_t) n_sys * sizeof(bfmg_batch_range_t);
	size_t const status_bytes = (size_t) n_sys * sizeof(bfmg_batch_status_t);

	void* ws = nullptr;

	if (bfmg_alloc(&ws, mat_bytes + 2 * vec_bytes + range_bytes + status_bytes + 512) < 0) {
		return -1;
	}

	char* at = (char*) ws;

	double* const scaled = (double*) at;
	at += mat_bytes;
	double* const dscale = (double*) at;
	at += vec_bytes;
	double* const bhat = (double*) at;
	at += vec_bytes;
	at = (char*) (((uintptr_t) at + 127) & ~(uintptr_t) 127);
	bfmg_batch_range_t* const d_ranges = (bfmg_batch_range_t*) at;
	at += range_bytes;
	at = (char*) (((uintptr_t) at + 127) & ~(uintptr_t) 127);
	bfmg_batch_status_t* const d_status = (bfmg_batch_status_t*) at;

	int rv = -1;
	int const t0 = bfmg_tick();

	size_t const smem = (size_t) max_rows * 4 * sizeof(double2);

	/* many systems: 256-thread CTAs, several per SM, so that independent systems hide each other's
	 * latencies; few systems: 1024 threads on each */

	bool const wide = n_sys < 2 * bfmg_sm_count();
	double2 const* const stop = (double2 const*) scaled;
	double2 const* const sbot = stop + pat->n_slots;

	if (
		BFMG_CHECK(cudaMemcpyAsync(d_ranges, ranges, range_bytes, cudaMemcpyHostToDevice, bfmg_stream())) < 0 ||
		bfmg_scale_system(pat, d_val, d_b, dscale, bhat, scaled) < 0
	) {
		goto out;
	}

	if (wide) {
		if (BFMG_CHECK(cudaFuncSetAttribute(k_pcg_cta<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) < 0) {
			goto out;
		}

		if (BFMG_LAUNCH(k_pcg_cta<1024>, n_sys, 1024, smem, *pat, stop, sbot, (double2 const*) bhat, (double2 const*) dscale, (double2*) d_x, d_ranges, d_status, opts->tol, opts->max_iter) < 0) {
			goto out;
		}
	}

	else {
		if (BFMG_CHECK(cudaFuncSetAttribute(k_pcg_cta<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) < 0) {
			goto out;
		}

		if (BFMG_LAUNCH(k_pcg_cta<256>, n_sys, 256, smem, *pat, stop, sbot, (double2 const*) bhat, (double2 const*) dscale, (double2*) d_x, d_ranges, d_status, opts->tol, opts->max_iter) < 0) {
			goto out;
		}
	}

	{
		int const t1 = bfmg_tick();

		if (bfmg_download(status, d_status, status_bytes) < 0) {
			goto out;
		}

		if (ms != nullptr) {
			*ms = bfmg_lap(t0, t1);
		}
	}

	rv = 0;

out:

	bfmg_free(ws);
	return rv;
}

} // extern "C"
