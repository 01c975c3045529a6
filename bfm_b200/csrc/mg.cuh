/*
 * mg.cuh - device side of the solver's multilevel preconditioner (included by solver.cu, after coarse.cuh).
 *
 * The reference's bfm_matrix_solve is a direct band LU (matrix.c:253-404, :531-541); here the system is solved
 * by conjugate gradients, and what decides the cost is the iteration count.  With a diagonal preconditioner it
 * grows like 1/h (46 000 iterations at 8 M DOF), with one additive coarse level of rigid-body modes like H/h
 * (1 520 at 50 M DOF, round 1).  This file replaces that level by an aggregation multigrid cycle on the
 * hierarchy of hier.c, used as a fixed symmetric positive definite preconditioner of plain PCG:
 *
 *   level 0   A^ = D^-1/2 A D^-1/2 (unit diagonal), SELL-32 2x2 node blocks - the assembled matrix
 *   level l   A_l = P^T A_{l-1} P, 3x3 node blocks (two translations + one rotation per aggregate), also
 *             scaled to a unit diagonal, SELL-32 with nine value planes
 *   last      dense, inverted once per solve by coarse.cuh's blocked Gauss-Jordan
 *
 * The tentative prolongator P~ holds, per node of an aggregate, the aggregate's rigid-body modes seen from that node,
 * in the scaled variables of both levels: P~_a = D_a^1/2 [1 0 -dy; 0 1 dx; (0 0 1)] D_I^-1/2.  With smoothed
 * aggregation (the default; hier.c lays the entries out, k_mg_smooth fills them) P = (I - w_s A^) P~ and the cycle is
 * a V-cycle; with P = P~ (BFM_MG_SMOOTH=0) it has to be a W-cycle.  One cycle on level l for a right-hand side g
 * (damped Jacobi, the diagonal being the identity):
 *
 *   z = w g                       pre-smoothing from a zero guess needs no product
 *   t = g - A z                   one fused SpMV (k_spmv_mg<kPre> / k_blk_spmv<kPre>)
 *   z += P cycle_{l+1}(P^T t)     restriction, recursion (gamma visits: V- or W-cycle), prolongation
 *   z += w (g - A z)              one fused SpMV (kPost); on level 0 it also leaves r.z for CG
 *
 * with w = 1.9 / (largest absolute row sum): a Gershgorin bound of the largest eigenvalue, so w lambda_max < 2
 * always holds and the smoother - hence the whole cycle - is symmetric positive definite.  The preconditioner
 * changes how fast CG converges, not what it converges to: the stopping test stays on the true CG residual.
 *
 * Set-up, once per solve: P from the diagonal scaling (and the level's scaled operator); A_{l+1} = P^T A_l P by
 * k_mg_ap + k_mg_ptq (Q = A P by fine node, P^T Q by coarse node) or, for one-entry rows, k_mg_rap (one warp per coarse
 * node, one pass over A_l); its diagonal, scaling and row-sum bound; at the end the dense inverse.  Everything is deterministic: fixed-order sums, no floating-point atomics (the row-sum bound
 * is a maximum, taken with an integer atomicMax on the bit pattern of non-negative doubles).
 *
 * All kernels are HBM-bound streaming work.  Algorithmic bytes per mesh node and PCG iteration on the
 * structured P1 plate (7 blocks per row, 2.65 prolongator entries per node), level 0:  k_spmv<kDot> 284 +
 * k_update_xr 96 + k_spmv_mg<kPre> 172 + k_mg_restrict 101 + k_mg_prolong 106 + k_spmv_mg<kPost> 188 + k_update_p 48
 * = 995 B; the coarser levels add about a tenth of that.
 */
#pragma once

#include <cstring>

namespace {

enum MgMode { kMgPlain, kMgPre, kMgResid, kMgPost };

constexpr int kMgGroup = 8;                       /* lanes that share one coarse node in k_mg_restrict */
constexpr int kMgGroupsPerBlock = kBlock / kMgGroup;

/* vector reads: through L1 (ld.global.nc) where a vector never changes during a launch; from L2 (ld.global.cg)
 * where a kernel reads what another SM - or a peer GPU - may have written while it runs */
template <bool CG>
__device__ __forceinline__ double ldv(double const* p) {
	return CG ? __ldcg(p) : __ldg(p);
}

struct MgDev {
	double omega[BFMG_MG_MAX_LEVELS];             /* damping of every level's Jacobi smoother */
	unsigned long long gersh[BFMG_MG_MAX_LEVELS]; /* bit pattern of the largest absolute row sum */
	int32_t bad;                                  /* a coarse operator came out with a non-positive diagonal */
	int32_t pad;
};

/* ---- level 0: fused SpMV variants over the SELL-32 2x2 node blocks ----------------------------------------
 *   kMgPre    out = g - w A^ g                 (v == g)
 *   kMgResid  out = g - A^ v
 *   kMgPost   out = v + w (g - A^ v),  + partial g.out; last CTA: rz -> beta, rho
 * FIRST: the preconditioner is applied to the initial residual (beta = 0, ignores S->done) */
template <MgMode MODE, bool FIRST>
__global__ void __launch_bounds__(kBlock) k_spmv_mg(
	const __grid_constant__ bfmg_pattern_t P, float2 const* __restrict__ vtop, float2 const* __restrict__ vbot,
	double2 const* __restrict__ v, double2 const* __restrict__ g, double2* __restrict__ out, double const* __restrict__ omega_p, double* __restrict__ partials, Scalars* S
) {
	pdl_sync();

	if (!FIRST && S->done) {
		return;
	}

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;
	double const w = MODE == kMgResid ? 1.0 : *omega_p;

	double acc = 0;

	for (int slice = P.row_lo / kWarp + warp; slice < (P.row_hi + kWarp - 1) / kWarp; slice += n_warps) {
		int const row = slice * kWarp + lane;
		int const beg = __ldg(&P.slice_off[slice]);
		int const end = __ldg(&P.slice_off[slice + 1]);

		double y0 = 0, y1 = 0;

#pragma unroll 8
		for (int slot = beg + lane; slot < end; slot += kWarp) {
			int const col = ld_stream(&P.scol[slot]);
			float2 const t = ld_stream(&vtop[slot]);
			float2 const u = ld_stream(&vbot[slot]);
			double2 const xv = __ldg(&v[col]);

			y0 = fma((double) t.x, xv.x, fma((double) t.y, xv.y, y0));
			y1 = fma((double) u.x, xv.x, fma((double) u.y, xv.y, y1));
		}

		if (row >= P.row_lo && row < P.row_hi) {
			double2 const gv = g[row];
			double2 o;

			if (MODE == kMgPost) {
				double2 const vv = __ldg(&v[row]);

				o.x = fma(w, gv.x - y0, vv.x);
				o.y = fma(w, gv.y - y1, vv.y);
				acc = fma(gv.x, o.x, fma(gv.y, o.y, acc));
			}

			else {
				o.x = fma(-w, y0, gv.x);
				o.y = fma(-w, y1, gv.y);
			}

			out[row] = o;
		}
	}

	if (MODE != kMgPost) {
		return;
	}

	double total;

	if (grid_sum(acc, partials, &S->ticket, &total) && threadIdx.x == 0) {
		if (S->world > 1) {
			S->part = total; /* this rank's share of r.z: k_share folds it together with r.r and x.x */
		}

		/* r.z of a symmetric positive definite preconditioner is positive; anything else is a breakdown */
		else if (!(total > 0) || isinf(total)) {
			S->done = S->done ? S->done : 2;
			S->beta = 0;
		}

		else {
			S->beta = FIRST ? 0 : total / S->rho;
			S->rho = total;
		}
	}
}

/* largest absolute row sum of the scaled level-0 operator, and its FP32 copy for the two smoothing products of the
 * cycle: the smoother only shapes the preconditioner - rounding its matrix to 24 bits perturbs M^-1 by ~1e-7, which
 * PCG does not notice, while the two products stream 20 instead of 36 bytes per block (CG's own product, the Galerkin
 * operators and the refinement residual use the FP64 values) */
__global__ void __launch_bounds__(kBlock) k_mg_gersh0(const __grid_constant__ bfmg_pattern_t P, double2 const* __restrict__ vtop, double2 const* __restrict__ vbot, float2* __restrict__ ftop, float2* __restrict__ fbot, unsigned long long* __restrict__ gersh) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;

	double worst = 0;

	for (int slice = P.row_lo / kWarp + warp; slice < (P.row_hi + kWarp - 1) / kWarp; slice += n_warps) {
		int const beg = __ldg(&P.slice_off[slice]);
		int const end = __ldg(&P.slice_off[slice + 1]);

		double s0 = 0, s1 = 0;

		for (int slot = beg + lane; slot < end; slot += kWarp) {
			double2 const t = ld_stream(&vtop[slot]);
			double2 const u = ld_stream(&vbot[slot]);

			ftop[slot] = make_float2((float) t.x, (float) t.y);
			fbot[slot] = make_float2((float) u.x, (float) u.y);

			s0 += fabs(t.x) + fabs(t.y);
			s1 += fabs(u.x) + fabs(u.y);
		}

		worst = fmax(worst, fmax(s0, s1));
	}

#pragma unroll
	for (int off = kWarp / 2; off > 0; off >>= 1) {
		worst = fmax(worst, __shfl_down_sync(0xffffffffu, worst, off));
	}

	if (lane == 0 && worst == worst) {
		atomicMax(gersh, (unsigned long long) __double_as_longlong(worst)); /* non-negative doubles order like their bits */
	}
}

/* several GPUs: all[r * stride + l] holds rank r's bound of level l (and, at index BFMG_MG_MAX_LEVELS, its "bad" flag):
 * every rank takes the maximum, so that all ranks smooth with the same damping and take the same decision */
__global__ void k_mg_omega(int n_levels, double factor, MgDev* D, double const* __restrict__ all, int world, int stride) {
	pdl_sync();

	int const l = threadIdx.x;

	if (l < n_levels) {
		double bound = __longlong_as_double((long long) D->gersh[l]);

		for (int r = 0; r < world; r++) {
			bound = fmax(bound, all[r * stride + l]);
		}

		D->gersh[l] = (unsigned long long) __double_as_longlong(bound);
		D->omega[l] = bound >= 1.0 ? factor / bound : factor; /* a unit diagonal makes every row sum >= 1 */
	}

	if (l == 0) {
		for (int r = 0; r < world; r++) {
			if (all[r * stride + BFMG_MG_MAX_LEVELS] != 0) {
				D->bad = 1;
			}
		}
	}
}

/* this rank's bounds and flag as doubles, for the all-gather */
__global__ void k_mg_bounds(MgDev const* D, int32_t const* bad, double* __restrict__ out) {
	pdl_sync();

	int const l = threadIdx.x;

	if (l < BFMG_MG_MAX_LEVELS) {
		out[l] = __longlong_as_double((long long) D->gersh[l]);
	}

	if (l == BFMG_MG_MAX_LEVELS) {
		out[l] = D->bad || *bad ? 1 : 0;
	}
}

/* several GPUs, first replicated level: every rank computed the rows of its own aggregates (the others are zero);
 * the operator is their sum, taken in rank order (exact: one non-zero contributor per entry) */
__global__ void k_mg_sum_ranks(size_t count, int world, double const* __restrict__ all, double* __restrict__ out) {
	pdl_sync();

	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < count; i += (size_t) gridDim.x * blockDim.x) {
		double t = 0;

		for (int r = 0; r < world; r++) {
			t += all[(size_t) r * count + i];
		}

		out[i] = t;
	}
}

/* ---- prolongator ------------------------------------------------------------------------------------------
 *
 * Values by entry, planes [k * 3 + m][entry] (k: unknown of the fine node, m: mode of the coarse node); PT is
 * float on level 0 (two thirds of the transfer traffic of an iteration is P, and its last digits only shape
 * the preconditioner - every level's operator is built from the SAME rounded P, so the cycle stays symmetric
 * positive definite), double above.  dsc = D^-1/2 of the FINE level (level 0: solver.cu's dscale). */
template <int NB, typename PT>
__global__ void k_mg_tentative(bfmg_mg_level_t L, double const* __restrict__ dsc, PT* __restrict__ pval) {
	pdl_sync();

	int const a = blockIdx.x * blockDim.x + threadIdx.x;

	if (a >= L.n) {
		return;
	}

	int const g = L.agg[a];
	float2 const d = ((float2 const*) L.geom)[a];
	size_t const np = (size_t) L.n_p;

	for (int e = L.p_ptr[a]; e < L.p_ptr[a + 1]; e++) {
		bool const own = L.p_col[e] == g;
		double s[3];

#pragma unroll
		for (int k = 0; k < NB; k++) {
			s[k] = own ? 1.0 / dsc[(size_t) NB * a + k] : 0.0;
		}

		/* rows of [1 0 -dy; 0 1 dx; 0 0 1] scaled by D_a^1/2 */
		pval[0 * np + e] = (PT) s[0];
		pval[1 * np + e] = (PT) 0;
		pval[2 * np + e] = (PT) (-(double) d.y * s[0]);
		pval[3 * np + e] = (PT) 0;
		pval[4 * np + e] = (PT) s[1];
		pval[5 * np + e] = (PT) ((double) d.x * s[1]);

		if (NB == 3) {
			pval[6 * np + e] = (PT) 0;
			pval[7 * np + e] = (PT) 0;
			pval[8 * np + e] = (PT) s[2];
		}
	}
}

/* P <- P D_c^-1/2 once the coarse level's scaling is known */
template <int NB, typename PT>
__global__ void k_mg_pscale(bfmg_mg_level_t L, double const* __restrict__ dsc_coarse, PT* __restrict__ pval) {
	pdl_sync();

	int const e = blockIdx.x * blockDim.x + threadIdx.x;

	if (e >= L.n_p) {
		return;
	}

	int const J = L.p_col[e];
	size_t const np = (size_t) L.n_p;

#pragma unroll
	for (int m = 0; m < 3; m++) {
		double const s = dsc_coarse[3 * (size_t) J + m];

#pragma unroll
		for (int k = 0; k < NB; k++) {
			pval[(size_t) (k * 3 + m) * np + e] = (PT) ((double) pval[(size_t) (k * 3 + m) * np + e] * s);
		}
	}
}

/* Smoothed aggregation: P <- (I - w A^) P~ with w = factor / (row-sum bound of A^), on the rows hier.c laid out for it
 * (build_transfer: nodes whose whole row is this rank's; everybody else keeps the tentative row).  One thread per
 * node; for each of its entries (coarse node J) the row of A^ is walked once and the columns b of aggregate J
 * contribute -w A^_ab P~_b, P~_b rebuilt from b's geometry and scaling - nothing is read from pval, so the update is in
 * place.  Fixed order: deterministic.  FINE as in k_mg_rap (level 0: the scaled operator's two planes of double2). */
template <int NB, typename PT>
__global__ void __launch_bounds__(kBlock) k_mg_smooth(bfmg_mg_level_t L, double const* __restrict__ fine, double const* __restrict__ dsc, PT* __restrict__ pval, unsigned long long const* __restrict__ gersh, double factor) {
	pdl_sync();

	int const a = L.row_lo + blockIdx.x * blockDim.x + threadIdx.x;

	if (a >= L.row_hi) {
		return;
	}

	int const ga = L.agg[a];
	int const abase = L.slice_off[a / kWarp] + a % kWarp;
	int const alen = L.row_len[a];

	if (ga < 0) {
		return;
	}

	for (int u = 0; u < alen; u++) { /* hier.c's rule: a ghost among the columns keeps the row tentative */
		int const b = L.scol[abase + u * kWarp];

		if (b < L.row_lo || b >= L.row_hi) {
			return;
		}
	}

	double const w = factor / __longlong_as_double((long long) *gersh);
	size_t const np = (size_t) L.n_p;
	size_t const ns = (size_t) L.n_slots;

	for (int e = L.p_ptr[a]; e < L.p_ptr[a + 1]; e++) {
		int const J = L.p_col[e];
		double acc[NB][3];

#pragma unroll
		for (int k = 0; k < NB; k++) {
#pragma unroll
			for (int m = 0; m < 3; m++) {
				acc[k][m] = 0;
			}
		}

		if (J == ga) {
			float2 const d = ((float2 const*) L.geom)[a];
			double const s0 = 1.0 / dsc[(size_t) NB * a + 0];
			double const s1 = 1.0 / dsc[(size_t) NB * a + 1];

			acc[0][0] = s0, acc[0][2] = -(double) d.y * s0;
			acc[1][1] = s1, acc[1][2] = (double) d.x * s1;

			if (NB == 3) {
				acc[NB - 1][2] = 1.0 / dsc[(size_t) NB * a + (NB - 1)];
			}
		}

		for (int u = 0; u < alen; u++) {
			int const slot = abase + u * kWarp;
			int const b = L.scol[slot];

			if (L.agg[b] != J) {
				continue;
			}

			double A[NB][NB];

			if (NB == 2) {
				double2 const top = ((double2 const*) fine)[slot];
				double2 const bot = ((double2 const*) fine)[ns + slot];

				A[0][0] = top.x, A[0][1] = top.y;
				A[1][0] = bot.x, A[1][1] = bot.y;
			}

			else {
#pragma unroll
				for (int k = 0; k < NB; k++) {
#pragma unroll
					for (int j = 0; j < NB; j++) {
						A[k][j] = fine[(size_t) (3 * k + j) * ns + slot];
					}
				}
			}

			/* P~_b: rows of [1 0 -dy; 0 1 dx; 0 0 1] scaled by D_b^1/2 (k_mg_tentative) */
			float2 const d = ((float2 const*) L.geom)[b];
			double pb[NB][3];

#pragma unroll
			for (int j = 0; j < NB; j++) {
#pragma unroll
				for (int m = 0; m < 3; m++) {
					pb[j][m] = 0;
				}
			}

			{
				double const s0 = 1.0 / dsc[(size_t) NB * b + 0];
				double const s1 = 1.0 / dsc[(size_t) NB * b + 1];

				pb[0][0] = s0, pb[0][2] = -(double) d.y * s0;
				pb[1][1] = s1, pb[1][2] = (double) d.x * s1;

				if (NB == 3) {
					pb[NB - 1][2] = 1.0 / dsc[(size_t) NB * b + (NB - 1)];
				}
			}

#pragma unroll
			for (int k = 0; k < NB; k++) {
#pragma unroll
				for (int m = 0; m < 3; m++) {
					double t = 0;

#pragma unroll
					for (int j = 0; j < NB; j++) {
						t = fma(A[k][j], pb[j][m], t);
					}

					acc[k][m] = fma(-w, t, acc[k][m]);
				}
			}
		}

#pragma unroll
		for (int k = 0; k < NB; k++) {
#pragma unroll
			for (int m = 0; m < 3; m++) {
				pval[(size_t) (k * 3 + m) * np + e] = (PT) acc[k][m];
			}
		}
	}
}

/* out = P^T v: kMgGroup lanes per coarse node walk its entry list (ascending fine nodes), then a fixed
 * shuffle tree - deterministic.  n_out >= 3 * n_coarse: the padding up to it is zeroed (dense level). */
template <int NB, typename PT, bool CG>
__device__ __forceinline__ void d_mg_restrict(bfmg_mg_level_t const& L, PT const* __restrict__ pval, double const* __restrict__ v, double* __restrict__ out, int n_out) {
	int const sub = threadIdx.x & (kMgGroup - 1);
	int const n_nodes_out = (n_out + 2) / 3;
	size_t const np = (size_t) L.n_p;

	/* the loop bound is the same for all lanes of a warp (its groups take consecutive coarse nodes), so the
	 * full-mask shuffles below are always executed by all 32 lanes */
	for (int base = blockIdx.x * kMgGroupsPerBlock + (threadIdx.x / kWarp) * (kWarp / kMgGroup); base < n_nodes_out; base += gridDim.x * kMgGroupsPerBlock) {
		int const I = base + (threadIdx.x & (kWarp - 1)) / kMgGroup;
		double s[3] = {0, 0, 0};

		if (I < L.n_coarse) {
			int const end = L.r_ptr[I + 1];

			for (int at = L.r_ptr[I] + sub; at < end; at += kMgGroup) {
				int const e = L.r_ent[at];
				int const a = L.r_node[at];

#pragma unroll
				for (int k = 0; k < NB; k++) {
					double const x = ldv<CG>(&v[(size_t) NB * a + k]);

#pragma unroll
					for (int m = 0; m < 3; m++) {
						s[m] = fma((double) pval[(size_t) (k * 3 + m) * np + e], x, s[m]);
					}
				}
			}
		}

#pragma unroll
		for (int off = kMgGroup / 2; off > 0; off >>= 1) {
#pragma unroll
			for (int m = 0; m < 3; m++) {
				s[m] += __shfl_down_sync(0xffffffffu, s[m], off, kMgGroup);
			}
		}

		if (sub == 0 && I < n_nodes_out) {
#pragma unroll
			for (int m = 0; m < 3; m++) {
				if (3 * I + m < n_out) {
					out[3 * (size_t) I + m] = s[m];
				}
			}
		}
	}
}

template <int NB, typename PT>
__global__ void __launch_bounds__(kBlock) k_mg_restrict(bfmg_mg_level_t L, PT const* __restrict__ pval, double const* __restrict__ v, double* __restrict__ out, int n_out, Scalars const* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	d_mg_restrict<NB, PT, false>(L, pval, v, out, n_out);
}

/* z = w g + P mu (first visit of the coarse level), or z += P mu (later visits of a W-cycle) */
template <int NB, typename PT, bool ADD, bool CG>
__device__ __forceinline__ void d_mg_prolong(bfmg_mg_level_t const& L, int row0, int n_rows, PT const* __restrict__ pval, double const* __restrict__ mu, double const* __restrict__ g, double* __restrict__ z, double w) {
	size_t const np = (size_t) L.n_p;

	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
		int const a = row0 + i;
		double acc[3] = {0, 0, 0};

		for (int e = L.p_ptr[a]; e < L.p_ptr[a + 1]; e++) {
			int const J = L.p_col[e];

#pragma unroll
			for (int m = 0; m < 3; m++) {
				double const x = ldv<CG>(&mu[3 * (size_t) J + m]);

#pragma unroll
				for (int k = 0; k < NB; k++) {
					acc[k] = fma((double) pval[(size_t) (k * 3 + m) * np + e], x, acc[k]);
				}
			}
		}

#pragma unroll
		for (int k = 0; k < NB; k++) {
			size_t const at = (size_t) NB * a + k;
			z[at] = ADD ? ldv<CG>(&z[at]) + acc[k] : fma(w, ldv<CG>(&g[at]), acc[k]);
		}
	}
}

template <int NB, typename PT, bool ADD>
__global__ void __launch_bounds__(kBlock) k_mg_prolong(bfmg_mg_level_t L, int row0, int n_rows, PT const* __restrict__ pval, double const* __restrict__ mu, double const* __restrict__ g, double* __restrict__ z, double const* __restrict__ omega_p, Scalars const* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	d_mg_prolong<NB, PT, ADD, false>(L, row0, n_rows, pval, mu, g, z, *omega_p);
}

/* Galerkin operator of the next level, A_c = P^T A P, directly: one warp per coarse node I, lane = slot of row I in
 * the next level's pattern (hier.c computed that pattern symbolically as the very same product).  The warp walks
 * the fine nodes a of column I of P in ascending order and, for each, the blocks (a, b) of A's row in slot order;
 * the lane whose coarse column J owns b adds P_a^T A_ab P_b.  Every entry is accumulated by one lane in a fixed
 * order: deterministic, no atomics.  FINE: the operator of the fine level (level 0: the two planes of 2x2 blocks
 * (a00,a01) / (a10,a11); above: nine planes).  DENSE: the result goes into the row-major n_dense x n_dense matrix. */
template <int NB, typename PT, bool DENSE>
__global__ void __launch_bounds__(kBlock) k_mg_rap(bfmg_mg_level_t L, bfmg_mg_level_t N, double const* __restrict__ fine, PT const* __restrict__ pval, double* __restrict__ val, int n_dense) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;
	size_t const np = (size_t) L.n_p;
	size_t const ns = (size_t) L.n_slots;

	for (int I = warp; I < N.row_hi; I += n_warps) { /* the rows this rank holds: all of them on one GPU */
		int const base = N.slice_off[I / kWarp] + I % kWarp;
		int const len = N.row_len[I];

		for (int t0 = 0; t0 < len; t0 += kWarp) { /* rows longer than a warp: one pass per 32 columns */
			int const my_slot = t0 + lane < len ? base + (t0 + lane) * kWarp : -1;
			int const my_col = my_slot >= 0 ? N.scol[my_slot] : -1;

			double c[3][3] = {};

			for (int at = L.r_ptr[I]; at < L.r_ptr[I + 1]; at++) {
				int const a = L.r_node[at];
				int const ea = L.r_ent[at];
				int const abase = L.slice_off[a / kWarp] + a % kWarp;
				int const alen = L.row_len[a];

				double pa[NB][3];

#pragma unroll
				for (int k = 0; k < NB; k++) {
#pragma unroll
					for (int m = 0; m < 3; m++) {
						pa[k][m] = (double) pval[(size_t) (k * 3 + m) * np + ea];
					}
				}

				for (int u = 0; u < alen; u++) {
					int const slot = abase + u * kWarp;
					int const b = L.scol[slot];
					int eb = L.p_ptr[b];
					int const eb_end = L.p_ptr[b + 1];

					for (; eb < eb_end && L.p_col[eb] != my_col; eb++) { /* one entry per node, or a few with a smoothed prolongator */
					}

					if (eb == eb_end) {
						continue; /* b is outside the coarse space, or has nothing for this lane's column */
					}

					double A[NB][NB];

					if (NB == 2) {
						double2 const top = ((double2 const*) fine)[slot];
						double2 const bot = ((double2 const*) fine)[ns + slot];

						A[0][0] = top.x, A[0][1] = top.y;
						A[1][0] = bot.x, A[1][1] = bot.y;
					}

					else {
#pragma unroll
						for (int k = 0; k < NB; k++) {
#pragma unroll
							for (int j = 0; j < NB; j++) {
								A[k][j] = fine[(size_t) (3 * k + j) * ns + slot];
							}
						}
					}

#pragma unroll
					for (int m = 0; m < 3; m++) {
						double t[NB];

#pragma unroll
						for (int k = 0; k < NB; k++) {
							t[k] = 0;

#pragma unroll
							for (int j = 0; j < NB; j++) {
								t[k] = fma(A[k][j], (double) pval[(size_t) (j * 3 + m) * np + eb], t[k]);
							}
						}

#pragma unroll
						for (int i = 0; i < 3; i++) {
#pragma unroll
							for (int k = 0; k < NB; k++) {
								c[i][m] = fma(pa[k][i], t[k], c[i][m]);
							}
						}
					}
				}
			}

			if (my_slot >= 0) {
#pragma unroll
				for (int i = 0; i < 3; i++) {
#pragma unroll
					for (int m = 0; m < 3; m++) {
						if (DENSE) {
							val[(size_t) (3 * I + i) * n_dense + 3 * my_col + m] = c[i][m];
						}

						else {
							val[(size_t) (3 * i + m) * N.n_slots + my_slot] = c[i][m];
						}
					}
				}
			}
		}
	}
}

/* The same product in two steps, for prolongators with several entries per node (smoothed aggregation).  Above, the
 * lane of column J visits every (a, b) pair of the row's support and looks b's entries up - with one entry per node a
 * third of the visits hit, with a smoothed prolongator (36 nodes in a support instead of 16, 2.7 entries per node, the
 * same 9 columns) one in seven: 92 ms instead of 33 on the mesh level at 50 M DOF.  So first
 *
 *   k_mg_ap    Q = A P by fine node: one thread per row a (SELL-32: lane = row, coalesced) accumulates
 *              Q(a, J) = sum_b A_ab P(b, J) over its slots in order, the J kept in a short list in local memory
 *              (first come, first listed; rap_width columns at most, else *overflow and the one-step kernel runs);
 *   k_mg_ptq   A_c(I, J) = sum_a P(a, I)^T Q(a, J): one warp per coarse node, lane = slot of row I as above, over the
 *              support's nodes in ascending order - 36 visits with a short search instead of 252.
 *
 * Every entry is still summed by one lane in a fixed order.  Q is n_own x rap_width blocks of NB x 3 doubles (+ the
 * column list): transient set-up memory, 15.6 GB for 25 M mesh nodes. */

template <int NB>
__host__ __device__ constexpr int rap_width() { return NB == 2 ? 12 : 24; } /* coarse columns a row of Q can hold: mesh level (aggregates of >= 3 nodes, 16 as a rule) / the levels above (6 nodes) */

template <int NB, typename PT>
__global__ void __launch_bounds__(kBlock) k_mg_ap(bfmg_mg_level_t L, double const* __restrict__ fine, PT const* __restrict__ pval, int32_t* __restrict__ qcol, double* __restrict__ qval, int32_t* __restrict__ overflow) {
	pdl_sync();

	int const a = blockIdx.x * blockDim.x + threadIdx.x; /* from row 0: lane = row inside the slice */

	if (a < L.row_lo || a >= L.row_hi) {
		return;
	}

	constexpr int kRapWidth = rap_width<NB>();

	int const abase = L.slice_off[a / kWarp] + a % kWarp;
	int const alen = L.row_len[a];
	size_t const np = (size_t) L.n_p;
	size_t const ns = (size_t) L.n_slots;

	int32_t col[kRapWidth];
	double q[kRapWidth][NB * 3];
	int cnt = 0;

	for (int u = 0; u < alen; u++) {
		int const slot = abase + u * kWarp;
		int const b = L.scol[slot];
		int const eb_end = L.p_ptr[b + 1];

		double A[NB][NB];

		if (NB == 2) {
			double2 const top = ((double2 const*) fine)[slot];
			double2 const bot = ((double2 const*) fine)[ns + slot];

			A[0][0] = top.x, A[0][1] = top.y;
			A[1][0] = bot.x, A[1][1] = bot.y;
		}

		else {
#pragma unroll
			for (int k = 0; k < NB; k++) {
#pragma unroll
				for (int j = 0; j < NB; j++) {
					A[k][j] = fine[(size_t) (3 * k + j) * ns + slot];
				}
			}
		}

		for (int eb = L.p_ptr[b]; eb < eb_end; eb++) {
			int const J = L.p_col[eb];
			int t = 0;

			for (; t < cnt && col[t] != J; t++) {
			}

			if (t == cnt) {
				if (cnt == kRapWidth) {
					*overflow = 1;
					continue;
				}

				col[cnt++] = J;

#pragma unroll
				for (int i = 0; i < NB * 3; i++) {
					q[t][i] = 0;
				}
			}

#pragma unroll
			for (int m = 0; m < 3; m++) {
#pragma unroll
				for (int k = 0; k < NB; k++) {
					double acc = q[t][k * 3 + m];

#pragma unroll
					for (int j = 0; j < NB; j++) {
						acc = fma(A[k][j], (double) pval[(size_t) (j * 3 + m) * np + eb], acc);
					}

					q[t][k * 3 + m] = acc;
				}
			}
		}
	}

	size_t const row = (size_t) (a - L.row_lo) * kRapWidth;

	for (int t = 0; t < kRapWidth; t++) {
		qcol[row + t] = t < cnt ? col[t] : -1;

		if (t < cnt) {
#pragma unroll
			for (int i = 0; i < NB * 3; i++) {
				qval[(row + t) * (NB * 3) + i] = q[t][i];
			}
		}
	}
}

/* G lanes per coarse row: 32, or 16 where the rows are short (the level above the mesh has 9 columns per row on a
 * plate: two rows per warp keep 18 lanes busy instead of 9).  No shuffles: the two halves of a warp are independent. */
template <int NB, typename PT, bool DENSE, int G>
__global__ void __launch_bounds__(kBlock) k_mg_ptq(bfmg_mg_level_t L, bfmg_mg_level_t N, PT const* __restrict__ pval, int32_t const* __restrict__ qcol, double const* __restrict__ qval, double* __restrict__ val, int n_dense) {
	pdl_sync();

	constexpr int kRapWidth = rap_width<NB>();

	int const lane = threadIdx.x & (G - 1);
	int const group = (blockIdx.x * blockDim.x + threadIdx.x) / G;
	int const n_groups = gridDim.x * blockDim.x / G;
	size_t const np = (size_t) L.n_p;

	for (int I = group; I < N.row_hi; I += n_groups) {
		int const base = N.slice_off[I / kWarp] + I % kWarp;
		int const len = N.row_len[I];

		for (int t0 = 0; t0 < len; t0 += G) {
			int const my_slot = t0 + lane < len ? base + (t0 + lane) * kWarp : -1;
			int const my_col = my_slot >= 0 ? N.scol[my_slot] : -2; /* -2: matches neither a column nor the padding of a list */

			double c[3][3] = {};

			for (int at = L.r_ptr[I]; at < L.r_ptr[I + 1]; at++) {
				size_t const row = (size_t) (L.r_node[at] - L.row_lo) * kRapWidth;
				int const ea = L.r_ent[at];
				int t = 0;

				for (; t < kRapWidth && qcol[row + t] != my_col; t++) {
				}

				if (t == kRapWidth) {
					continue;
				}

				double const* const q = qval + (row + t) * (NB * 3);

#pragma unroll
				for (int i = 0; i < 3; i++) {
#pragma unroll
					for (int k = 0; k < NB; k++) {
						double const pa = (double) pval[(size_t) (k * 3 + i) * np + ea];

#pragma unroll
						for (int m = 0; m < 3; m++) {
							c[i][m] = fma(pa, q[k * 3 + m], c[i][m]);
						}
					}
				}
			}

			if (my_slot >= 0) {
#pragma unroll
				for (int i = 0; i < 3; i++) {
#pragma unroll
					for (int m = 0; m < 3; m++) {
						if (DENSE) {
							val[(size_t) (3 * I + i) * n_dense + 3 * my_col + m] = c[i][m];
						}

						else {
							val[(size_t) (3 * i + m) * N.n_slots + my_slot] = c[i][m];
						}
					}
				}
			}
		}
	}
}

/* ---- levels >= 1: 3x3 node blocks, SELL-32, value planes [3 k + m][slot] ---------------------------------- */

/* dsc = 1 / sqrt(diagonal); a diagonal that is not positive marks the hierarchy unusable */
__global__ void k_blk_diag(bfmg_mg_level_t N, double const* __restrict__ val, double* __restrict__ dsc, MgDev* D) {
	pdl_sync();

	int const I = blockIdx.x * blockDim.x + threadIdx.x;

	if (I >= N.row_hi) {
		return;
	}

	int const slot = N.diag_pos[I];

#pragma unroll
	for (int k = 0; k < 3; k++) {
		double const d = val[(size_t) (4 * k) * N.n_slots + slot];

		if (!(d > 0) || isinf(d)) {
			D->bad = 1;
			dsc[3 * (size_t) I + k] = 1;
		}

		else {
			dsc[3 * (size_t) I + k] = 1.0 / sqrt(d);
		}
	}
}

/* A <- D^-1/2 A D^-1/2 in place, its FP32 copy for the cycle's products (as on level 0: the cycle only shapes the
 * preconditioner; the Galerkin product of the next level reads the FP64 values), and the largest absolute row sum */
__global__ void __launch_bounds__(kBlock) k_blk_scale(bfmg_mg_level_t N, double* __restrict__ val, float* __restrict__ fval, double const* __restrict__ dsc, unsigned long long* __restrict__ gersh) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;

	double worst = 0;

	for (int slice = warp; slice < N.n_slices; slice += n_warps) {
		int const row = slice * kWarp + lane;
		int const end = N.slice_off[slice + 1];

		double sr[3] = {0, 0, 0};
		double sum[3] = {0, 0, 0};

		if (row < N.row_hi) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				sr[k] = dsc[3 * (size_t) row + k];
			}
		}

		for (int slot = N.slice_off[slice] + lane; slot < end; slot += kWarp) {
			int const col = N.scol[slot];

#pragma unroll
			for (int m = 0; m < 3; m++) {
				double const sc = dsc[3 * (size_t) col + m];

#pragma unroll
				for (int k = 0; k < 3; k++) {
					size_t const at = (size_t) (3 * k + m) * N.n_slots + slot;
					double const x = val[at] * sr[k] * sc;

					val[at] = x;
					fval[at] = (float) x;
					sum[k] += fabs(x);
				}
			}
		}

		worst = fmax(worst, fmax(sum[0], fmax(sum[1], sum[2])));
	}

#pragma unroll
	for (int off = kWarp / 2; off > 0; off >>= 1) {
		worst = fmax(worst, __shfl_down_sync(0xffffffffu, worst, off));
	}

	if (lane == 0 && worst == worst) {
		atomicMax(gersh, (unsigned long long) __double_as_longlong(worst));
	}
}

/* warp = slice, lane = node row; modes as k_spmv_mg (no dot product: the coarse levels feed no CG scalar) */
template <MgMode MODE, bool CG, typename VT>
__device__ __forceinline__ void d_blk_spmv(bfmg_mg_level_t const& N, VT const* __restrict__ val, double const* __restrict__ v, double const* __restrict__ g, double* __restrict__ out, double w) {
	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;
	size_t const ns = (size_t) N.n_slots;

	for (int slice = warp; slice < N.n_slices; slice += n_warps) {
		int const row = slice * kWarp + lane;
		int const end = __ldg(&N.slice_off[slice + 1]);

		double y0 = 0, y1 = 0, y2 = 0;

#pragma unroll 4
		for (int slot = __ldg(&N.slice_off[slice]) + lane; slot < end; slot += kWarp) {
			int const col = ld_stream(&N.scol[slot]);
			double const x0 = ldv<CG>(&v[3 * (size_t) col + 0]);
			double const x1 = ldv<CG>(&v[3 * (size_t) col + 1]);
			double const x2 = ldv<CG>(&v[3 * (size_t) col + 2]);

			y0 = fma((double) ld_stream(&val[0 * ns + slot]), x0, fma((double) ld_stream(&val[1 * ns + slot]), x1, fma((double) ld_stream(&val[2 * ns + slot]), x2, y0)));
			y1 = fma((double) ld_stream(&val[3 * ns + slot]), x0, fma((double) ld_stream(&val[4 * ns + slot]), x1, fma((double) ld_stream(&val[5 * ns + slot]), x2, y1)));
			y2 = fma((double) ld_stream(&val[6 * ns + slot]), x0, fma((double) ld_stream(&val[7 * ns + slot]), x1, fma((double) ld_stream(&val[8 * ns + slot]), x2, y2)));
		}

		if (row < N.row_hi) {
			size_t const at = 3 * (size_t) row;
			double const y[3] = {y0, y1, y2};

#pragma unroll
			for (int k = 0; k < 3; k++) {
				double o;

				if (MODE == kMgPlain) {
					o = y[k];
				}

				else if (MODE == kMgPost) {
					o = fma(w, ldv<CG>(&g[at + k]) - y[k], ldv<CG>(&v[at + k]));
				}

				else {
					o = fma(-w, y[k], ldv<CG>(&g[at + k]));
				}

				out[at + k] = o;
			}
		}
	}
}

template <MgMode MODE, typename VT>
__global__ void __launch_bounds__(kBlock) k_blk_spmv(bfmg_mg_level_t N, VT const* __restrict__ val, double const* __restrict__ v, double const* __restrict__ g, double* __restrict__ out, double const* __restrict__ omega_p, Scalars const* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	d_blk_spmv<MODE, false, VT>(N, val, v, g, out, (MODE == kMgPre || MODE == kMgPost) ? *omega_p : 1.0);
}

/* mu = E^-1 g on the dense last level (the explicit inverse; kCoarseRows rows per CTA, four warps per row as in
 * k_coarse_apply) */
__device__ __forceinline__ void d_dense_apply(int nc, double const* __restrict__ Einv, double const* __restrict__ g, double* __restrict__ mu) {
	__shared__ double quarter_sum[kWarpsPerBlock];

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = threadIdx.x / kWarp;
	int const span = ((nc + 3) / 4 + kWarp - 1) / kWarp * kWarp;
	int const j_end = min(nc, (warp % 4 + 1) * span);
	int const n_row_blocks = (nc + kCoarseRows - 1) / kCoarseRows;

	for (int rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x) {
		int const row = rb * kCoarseRows + warp / 4;
		double t = 0;

		if (row < nc) {
			double const* const e = Einv + (size_t) row * nc;
			double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
			int j = (warp % 4) * span + lane;

			for (; j + 3 * kWarp < j_end; j += 4 * kWarp) {
				t0 = fma(ld_stream(&e[j]), __ldcg(&g[j]), t0);
				t1 = fma(ld_stream(&e[j + kWarp]), __ldcg(&g[j + kWarp]), t1);
				t2 = fma(ld_stream(&e[j + 2 * kWarp]), __ldcg(&g[j + 2 * kWarp]), t2);
				t3 = fma(ld_stream(&e[j + 3 * kWarp]), __ldcg(&g[j + 3 * kWarp]), t3);
			}

			for (; j < j_end; j += kWarp) {
				t0 = fma(ld_stream(&e[j]), __ldcg(&g[j]), t0);
			}

			t = warp_sum((t0 + t1) + (t2 + t3));
		}

		__syncthreads(); /* the previous row block's sums have been consumed */

		if (lane == 0) {
			quarter_sum[warp] = t;
		}

		__syncthreads();

		if (lane == 0 && warp % 4 == 0 && row < nc) {
			mu[row] = (quarter_sum[warp] + quarter_sum[warp + 1]) + (quarter_sum[warp + 2] + quarter_sum[warp + 3]);
		}
	}
}

__global__ void __launch_bounds__(kBlock) k_dense_apply(int nc, double const* __restrict__ Einv, double const* __restrict__ g, double* __restrict__ mu, Scalars const* S, bool obey_done) {
	pdl_sync();

	if (obey_done && S->done) {
		return;
	}

	d_dense_apply(nc, Einv, g, mu);
}

/* identity on the padding rows of the dense operator (3 n .. nc) */
__global__ void k_mg_dense_pad(int n_real, int nc, double* __restrict__ E) {
	pdl_sync();

	int const i = n_real + blockIdx.x * blockDim.x + threadIdx.x;

	if (i < nc) {
		E[(size_t) i * nc + i] = 1;
	}
}

/* ---- one solve's multilevel state -------------------------------------------------------------------------- */

struct MgLevelWork {
	bfmg_mg_level_t L;
	double* val;      /* levels >= 1 (sparse): nine planes of n_slots doubles */
	float* fval;      /* ... and their FP32 copy, which the cycle's products stream */
	double* dsc;      /* levels >= 1: D^-1/2, three per node */
	void* pval;       /* prolongator to the next level: float [6][n_p] on level 0, double [9][n_p] above */
	double *g, *va, *vb, *vt; /* levels >= 1: right-hand side and three work vectors, three per node */
	int gamma;        /* visits of the next level per cycle */
	int grid_rows;    /* grid of the per-slice kernels */
	int grid_nodes;   /* grid of the per-node kernels */
};

struct MgRun {
	bfmg_mg_t const* M = nullptr;
	int n_levels = 0;
	int nc = 0;
	void* ws = nullptr;
	MgDev* D = nullptr;
	MgLevelWork W[BFMG_MG_MAX_LEVELS] = {};

	/* dense last level */
	double* E = nullptr;
	double* Pblk = nullptr;
	double* dg = nullptr;   /* its right-hand side [nc] */
	double* dmu = nullptr;  /* its solution [nc] */
	int32_t* bad = nullptr;

	double omega_factor = 1.9;        /* Jacobi damping: w = omega_factor / (row-sum bound) < 2 / lambda_max whatever the matrix (BFM_MG_OMEGA);
	                                   * 1.8 -> 1.9: 72 -> 67 iterations at 50 M DOF */
	bool rap_two_step = true;         /* BFM_MG_RAP=direct: the one-step Galerkin kernel for smoothed prolongators too */
	double smooth_factor = 1.8;       /* prolongator smoothing: w = smooth_factor / (row-sum bound), BFM_MG_SMOOTH_OMEGA.  The bound is
	                                   * Gershgorin's, ~1.4x the largest eigenvalue on these operators: 1.8 / bound ~ the classic 4 / (3 lambda_max) */
	double lambda0 = 4;     /* Gershgorin bound of the scaled level-0 operator (after setup) */
	float2* ftop = nullptr; /* FP32 copy of the scaled level-0 operator for the smoothing products: plane of (a00,a01) */
	float2* fbot = nullptr; /* ... and of (a10,a11) */

	/* several GPUs (exchanges over NVLink peer memory, p2p.cuh): the levels 0 .. first_rep - 1 are distributed - a halo
	 * exchange before every product with their operators - the levels from first_rep on are held by every rank */
	bool dist = false;
	int world = 1;
	int first_rep = 0;
	HaloDev HD[BFMG_MG_MAX_LEVELS] = {};
	double* gpart = nullptr;   /* restriction onto the first replicated level: this rank's entries, the others zero */
	double* gath = nullptr;    /* set-up: every rank's partial operator of the first replicated level; bounds of all ranks */
	size_t gath_count = 0;     /* doubles of that operator (9 n_slots, or nc^2 when it is the dense level) */
	double* bounds = nullptr;  /* [world + 1][BFMG_MG_MAX_LEVELS + 1] */

	static size_t align256(size_t v) { return (v + 255) & ~(size_t) 255; }

	bool rep_is_dense() const { return first_rep == n_levels - 1; }

	int alloc(bfmg_mg_t const* mg, int world_) {
		M = mg;
		n_levels = mg->n_levels;
		nc = mg->nc;
		world = world_;
		dist = world > 1;
		first_rep = 0;

		while (dist && first_rep < n_levels && mg->level[first_rep].distributed) {
			first_rep++;
		}

		if (dist && (first_rep == 0 || first_rep >= n_levels)) {
			bfmg_set_error("multigrid hierarchy of a partitioned job has no replicated level");
			return -1;
		}

		size_t total = align256(sizeof(MgDev)) + 2 * align256((size_t) mg->level[0].n_slots * sizeof(float2));

		for (int l = 0; l < n_levels; l++) {
			bfmg_mg_level_t const& L = mg->level[l];
			bool const last = l == n_levels - 1;

			if (l >= 1 && !last) {
				total += align256((size_t) L.n_slots * 9 * sizeof(double)) + align256((size_t) L.n_slots * 9 * sizeof(float));
			}

			if (l >= 1) {
				total += align256((size_t) L.n * 3 * sizeof(double));      /* dsc */
				total += 4 * align256((size_t) L.n * 3 * sizeof(double));  /* g, va, vb, vt */
			}

			if (!last) {
				total += align256((size_t) L.n_p * (l == 0 ? 6 * sizeof(float) : 9 * sizeof(double)));
			}
		}

		total += align256((size_t) nc * nc * sizeof(double)) + align256(kGjBlock * kGjBlock * sizeof(double)) + 2 * align256(((size_t) nc + 8) * sizeof(double)) + 256;

		if (dist) {
			gath_count = rep_is_dense() ? (size_t) nc * nc : (size_t) mg->level[first_rep].n_slots * 9;
			total += align256(((size_t) next_len(first_rep - 1) + 8) * sizeof(double));
			total += align256((size_t) world * gath_count * sizeof(double));
			total += align256((size_t) (world + 1) * (BFMG_MG_MAX_LEVELS + 1) * sizeof(double));
		}

		if (bfmg_alloc(&ws, total) < 0) {
			return -1;
		}

		char* at = (char*) ws;
		auto take = [&](size_t bytes) { char* const p = at; at += align256(bytes); return (void*) p; };

		D = (MgDev*) take(sizeof(MgDev));
		ftop = (float2*) take((size_t) mg->level[0].n_slots * sizeof(float2));
		fbot = (float2*) take((size_t) mg->level[0].n_slots * sizeof(float2));

		for (int l = 0; l < n_levels; l++) {
			bfmg_mg_level_t const& L = mg->level[l];
			bool const last = l == n_levels - 1;
			MgLevelWork& w = W[l];

			w.L = L;
			w.gamma = 1;
			w.grid_rows = bfmg_grid((L.n_slices + kWarpsPerBlock - 1) / kWarpsPerBlock, 8);
			w.grid_nodes = bfmg_grid(((int64_t) L.n + kBlock - 1) / kBlock, 8);

			if (l >= 1 && !last) {
				w.val = (double*) take((size_t) L.n_slots * 9 * sizeof(double));
				w.fval = (float*) take((size_t) L.n_slots * 9 * sizeof(float));
			}

			if (l >= 1) {
				w.dsc = (double*) take((size_t) L.n * 3 * sizeof(double));
				w.g = (double*) take((size_t) L.n * 3 * sizeof(double));
				w.va = (double*) take((size_t) L.n * 3 * sizeof(double));
				w.vb = (double*) take((size_t) L.n * 3 * sizeof(double));
				w.vt = (double*) take((size_t) L.n * 3 * sizeof(double));
			}

			if (!last) {
				w.pval = take((size_t) L.n_p * (l == 0 ? 6 * sizeof(float) : 9 * sizeof(double)));
			}

			HaloDev& H = HD[l];

			H.n_nbr = L.distributed ? L.n_nbr : 0;
			H.n_send = L.distributed ? L.n_send : 0;

			for (int k = 0; k < H.n_nbr; k++) {
				H.nbr[k] = L.nbr[k];
				H.recv_begin[k] = L.recv_begin[k];
				H.recv_count[k] = L.recv_count[k];
				H.send_ptr[k] = L.send_ptr[k];
				H.send_ptr[k + 1] = L.send_ptr[k + 1];
			}
		}

		E = (double*) take((size_t) nc * nc * sizeof(double));
		Pblk = (double*) take(kGjBlock * kGjBlock * sizeof(double));
		dg = (double*) take(((size_t) nc + 8) * sizeof(double));
		dmu = (double*) take(((size_t) nc + 8) * sizeof(double));
		bad = (int32_t*) take(256);

		if (dist) {
			gpart = (double*) take(((size_t) next_len(first_rep - 1) + 8) * sizeof(double));
			gath = (double*) take((size_t) world * gath_count * sizeof(double));
			bounds = (double*) take((size_t) (world + 1) * (BFMG_MG_MAX_LEVELS + 1) * sizeof(double));
		}

		/* cycle shape: BFM_MG_GAMMA visits of the next level on every sparse level above the mesh (W-cycle by
		 * default: the piecewise-rigid interpolation of plain aggregation needs it to stay level-independent) */

		char const* env = getenv("BFM_MG_GAMMA");

		for (int l = 1; l < n_levels - 1; l++) { /* "2" or a list per level from level 1 on, "2,2,1": the last entry repeats */
			int g = mg->level[l].smoothed ? 1 : 2; /* a smoothed prolongator interpolates well enough for the V-cycle */

			if (env != nullptr && env[0] != 0) {
				char const* at_env = env;

				for (int skip = 1; skip < l; skip++) {
					char const* const comma = strchr(at_env, ',');

					if (comma == nullptr) {
						break;
					}

					at_env = comma + 1;
				}

				g = atoi(at_env);
			}

			W[l].gamma = g >= 1 && g <= 4 ? g : 2;
		}

		env = getenv("BFM_MG_OMEGA");

		if (env != nullptr && atof(env) > 0 && atof(env) < 2) {
			omega_factor = atof(env);
		}

		env = getenv("BFM_MG_RAP");
		rap_two_step = env == nullptr || strcmp(env, "direct") != 0;

		env = getenv("BFM_MG_SMOOTH_OMEGA");

		if (env != nullptr && atof(env) > 0 && atof(env) < 2) {
			smooth_factor = atof(env);
		}

		return 0;
	}

	/* several GPUs: do the halos of every distributed level and the gathered vector fit the mailboxes? */
	bool fits(P2p const* px) const {
		if (!dist) {
			return true;
		}

		if (px == nullptr || next_len(first_rep - 1) > px->L.gather_cap) {
			return false;
		}

		for (int l = 0; l < first_rep; l++) {
			int const nb = l == 0 ? 2 : 3;

			if (HD[l].n_nbr > kP2pMaxRanks) {
				return false;
			}

			for (int k = 0; k < HD[l].n_nbr; k++) {
				if ((int64_t) HD[l].recv_count[k] * nb > (int64_t) px->L.halo_cap * 2 || (int64_t) (HD[l].send_ptr[k + 1] - HD[l].send_ptr[k]) * nb > (int64_t) px->L.halo_cap * 2) {
					return false;
				}
			}
		}

		return true;
	}

	void release() {
		bfmg_free(ws);
		ws = nullptr;
	}

	/* refresh the ghosts of a vector of the distributed level l from their owners (nothing elsewhere) */
	bool halo(int l, double* vec, Scalars* S, bool obey) {
		if (!dist || l >= first_rep) {
			return true;
		}

		HaloDev const& H = HD[l];
		int const grid = H.n_send > 0 ? (H.n_send + kBlock - 1) / kBlock : 1;

		int rc = l == 0
			? BFMG_LAUNCH(k_halo_post<2>, grid, kBlock, 0, H, (double const*) vec, (int32_t const*) W[l].L.send_idx, S, obey)
			: BFMG_LAUNCH(k_halo_post<3>, grid, kBlock, 0, H, (double const*) vec, (int32_t const*) W[l].L.send_idx, S, obey);

		if (rc == 0 && H.n_nbr > 0) {
			rc = l == 0
				? BFMG_LAUNCH(k_halo_take<2>, H.n_nbr, kBlock, 0, H, vec, S, obey)
				: BFMG_LAUNCH(k_halo_take<3>, H.n_nbr, kBlock, 0, H, vec, S, obey);
		}

		return rc == 0;
	}

	/* out = P_l^T v */
	bool restrict_to(int l, double const* v, double* out, int n_out, Scalars* S, bool obey) {
		MgLevelWork const& w = W[l];
		int const grid = bfmg_grid(((int64_t) (n_out + 2) / 3 + kMgGroupsPerBlock - 1) / kMgGroupsPerBlock, 8);

		return l == 0
			? BFMG_LAUNCH((k_mg_restrict<2, float>), grid, kBlock, 0, w.L, (float const*) w.pval, v, out, n_out, S, obey) == 0
			: BFMG_LAUNCH((k_mg_restrict<3, double>), grid, kBlock, 0, w.L, (double const*) w.pval, v, out, n_out, S, obey) == 0;
	}

	/* right-hand side / solution buffers of level l + 1 as seen from level l */
	double* next_g(int l) { return l + 1 == n_levels - 1 ? dg : W[l + 1].g; }
	int next_len(int l) const { return l + 1 == n_levels - 1 ? nc : 3 * M->level[l + 1].n; }

	/* right-hand side of level l + 1 = P_l^T v.  A distributed level produces the entries of its own aggregates: the
	 * owned part of a distributed next level (its ghosts follow by halo()), or - below the first replicated level - a
	 * slice that every rank stores into every mailbox, from where the complete vector is taken */
	bool restrict_down(int l, double const* v, Scalars* S, bool obey) {
		if (dist && l + 1 == first_rep) {
			int const real = 3 * M->level[l + 1].n;

			return
				restrict_to(l, v, gpart, real, S, obey) &&
				BFMG_LAUNCH(k_mg_gather_post, bfmg_grid(((int64_t) 3 * W[l].L.gather_count + kBlock - 1) / kBlock, 2), kBlock, 0, 3 * W[l].L.gather_first, 3 * W[l].L.gather_count, (double const*) gpart, S, obey) == 0 &&
				BFMG_LAUNCH(k_mg_gather_take, bfmg_grid(((int64_t) real + kBlock - 1) / kBlock, 2), kBlock, 0, real, next_g(l), S, obey) == 0;
		}

		if (dist && l + 1 < first_rep) {
			return restrict_to(l, v, next_g(l), 3 * M->level[l + 1].row_hi, S, obey);
		}

		return restrict_to(l, v, next_g(l), next_len(l), S, obey);
	}

	/* set-up: prolongators, coarse operators (Galerkin products), scalings, damping factors, the dense inverse.
	 * *usable = false when a coarse operator is not positive definite (the caller then solves with the diagonal
	 * preconditioner alone; on several GPUs every rank takes the same decision) */
	int setup(bfmg_pattern_t const* pat, double2 const* stop, double2 const* sbot, double2 const* dscale, int spmv_grid, Scalars* S, bool* usable) {
		*usable = false;

		if (
			BFMG_CHECK(cudaMemsetAsync(D, 0, sizeof(MgDev), bfmg_stream())) < 0 ||
			BFMG_CHECK(cudaMemsetAsync(bad, 0, sizeof(int32_t), bfmg_stream())) < 0 ||
			BFMG_CHECK(cudaMemsetAsync(dg, 0, ((size_t) nc + 8) * sizeof(double), bfmg_stream())) < 0 ||
			BFMG_LAUNCH(k_mg_gersh0, spmv_grid, kBlock, 0, *pat, stop, sbot, ftop, fbot, &D->gersh[0]) < 0
		) {
			return -1;
		}

		for (int l = 0; l + 1 < n_levels; l++) {
			MgLevelWork& w = W[l];
			MgLevelWork& nx = W[l + 1];
			bool const dense = l + 1 == n_levels - 1;
			bool const transition = dist && l + 1 == first_rep; /* this level is distributed, the next one replicated */
			int const node_blocks = (w.L.n + kBlock - 1) / kBlock;
			int const coarse_blocks = (nx.L.row_hi + kBlock - 1) / kBlock;
			double* const target = dense ? E : nx.val;
			size_t const target_bytes = dense ? (size_t) nc * nc * sizeof(double) : (size_t) nx.L.n_slots * 9 * sizeof(double);

			int rc = l == 0
				? BFMG_LAUNCH((k_mg_tentative<2, float>), node_blocks, kBlock, 0, w.L, (double const*) dscale, (float*) w.pval)
				: BFMG_LAUNCH((k_mg_tentative<3, double>), node_blocks, kBlock, 0, w.L, (double const*) w.dsc, (double*) w.pval);

			if (rc == 0 && w.L.smoothed) { /* smoothed aggregation: (I - w A^) P~ with the scaled operator of this level */
				int const own_blocks = (w.L.row_hi - w.L.row_lo + kBlock - 1) / kBlock;

				rc = l == 0
					? BFMG_LAUNCH((k_mg_smooth<2, float>), own_blocks, kBlock, 0, w.L, (double const*) stop, (double const*) dscale, (float*) w.pval, (unsigned long long const*) &D->gersh[0], smooth_factor)
					: BFMG_LAUNCH((k_mg_smooth<3, double>), own_blocks, kBlock, 0, w.L, (double const*) w.val, (double const*) w.dsc, (double*) w.pval, (unsigned long long const*) &D->gersh[l], smooth_factor);
			}

			if (rc < 0 || BFMG_CHECK(cudaMemsetAsync(target, 0, target_bytes, bfmg_stream())) < 0) {
				return -1;
			}

			int const rap_grid = bfmg_grid(((int64_t) nx.L.row_hi + kWarpsPerBlock - 1) / kWarpsPerBlock, 8);

			/* a prolongator with several entries per node: Q = A P by fine node, then P^T Q (see k_mg_ap); the one-step
			 * kernel when a node's Q row does not fit kRapWidth columns, or without the memory for Q */

			bool two_step = w.L.smoothed != 0 && rap_two_step;

			if (two_step) {
				size_t const n_own = (size_t) (w.L.row_hi - w.L.row_lo);
				int const nb = l == 0 ? 2 : 3;
				size_t const width = (size_t) (l == 0 ? rap_width<2>() : rap_width<3>());
				int32_t* qcol = nullptr;
				double* qval = nullptr;
				int32_t overflow = 0;

				if (bfmg_alloc((void**) &qcol, n_own * width * sizeof *qcol) < 0 || bfmg_alloc((void**) &qval, n_own * width * nb * 3 * sizeof *qval) < 0) {
					two_step = false;
				}

				else {
					int const ap_grid = (w.L.row_hi + kBlock - 1) / kBlock;

					rc = BFMG_CHECK(cudaMemsetAsync(bad + 8, 0, sizeof(int32_t), bfmg_stream()));

					rc = rc < 0 ? rc : (l == 0
						? BFMG_LAUNCH((k_mg_ap<2, float>), ap_grid, kBlock, 0, w.L, (double const*) stop, (float const*) w.pval, qcol, qval, bad + 8)
						: BFMG_LAUNCH((k_mg_ap<3, double>), ap_grid, kBlock, 0, w.L, (double const*) w.val, (double const*) w.pval, qcol, qval, bad + 8));

					if (rc == 0) {
						int const half_grid = bfmg_grid(((int64_t) nx.L.row_hi + 2 * kWarpsPerBlock - 1) / (2 * kWarpsPerBlock), 8); /* two rows per warp */

						rc = dense
							? (l == 0
								? BFMG_LAUNCH((k_mg_ptq<2, float, true, 32>), rap_grid, kBlock, 0, w.L, nx.L, (float const*) w.pval, (int32_t const*) qcol, (double const*) qval, target, nc)
								: BFMG_LAUNCH((k_mg_ptq<3, double, true, 32>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) w.pval, (int32_t const*) qcol, (double const*) qval, target, nc))
							: (l == 0
								? BFMG_LAUNCH((k_mg_ptq<2, float, false, 16>), half_grid, kBlock, 0, w.L, nx.L, (float const*) w.pval, (int32_t const*) qcol, (double const*) qval, target, nc)
								: BFMG_LAUNCH((k_mg_ptq<3, double, false, 32>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) w.pval, (int32_t const*) qcol, (double const*) qval, target, nc));
					}

					if (rc == 0 && (BFMG_CHECK(cudaMemcpyAsync(&overflow, bad + 8, sizeof overflow, cudaMemcpyDeviceToHost, bfmg_stream())) < 0 || BFMG_CHECK(cudaStreamSynchronize(bfmg_stream())) < 0)) {
						rc = -1;
					}

					two_step = overflow == 0;
				}

				bfmg_free(qcol);
				bfmg_free(qval);

				if (rc < 0) {
					return -1;
				}
			}

			if (two_step) {
				/* done above */
			}

			else if (dense) {
				rc = l == 0
					? BFMG_LAUNCH((k_mg_rap<2, float, true>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) stop, (float const*) w.pval, target, nc)
					: BFMG_LAUNCH((k_mg_rap<3, double, true>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) w.val, (double const*) w.pval, target, nc);
			}

			else {
				rc = l == 0
					? BFMG_LAUNCH((k_mg_rap<2, float, false>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) stop, (float const*) w.pval, target, nc)
					: BFMG_LAUNCH((k_mg_rap<3, double, false>), rap_grid, kBlock, 0, w.L, nx.L, (double const*) w.val, (double const*) w.pval, target, nc);
			}

			if (rc < 0) {
				return -1;
			}

			if (transition) {
				/* every rank holds the rows of its own aggregates: add the ranks' parts up (one collective, set-up only) */

				if (
					bfmg_dist_allgather_f64(target, gath, (int) gath_count) < 0 ||
					BFMG_LAUNCH(k_mg_sum_ranks, bfmg_grid(((int64_t) gath_count + kBlock - 1) / kBlock, 8), kBlock, 0, gath_count, world, (double const*) gath, target) < 0
				) {
					return -1;
				}
			}

			if (dense) {
				if (3 * nx.L.n < nc && BFMG_LAUNCH(k_mg_dense_pad, 1, kBlock, 0, 3 * nx.L.n, nc, E) < 0) {
					return -1;
				}
			}

			else {
				rc = BFMG_LAUNCH(k_blk_diag, coarse_blocks, kBlock, 0, nx.L, nx.val, nx.dsc, D);

				if (rc == 0 && !halo(l + 1, nx.dsc, S, false)) { /* the scaling of the ghost columns */
					rc = -1;
				}

				rc = rc < 0 ? rc : BFMG_LAUNCH(k_blk_scale, nx.grid_rows, kBlock, 0, nx.L, nx.val, nx.fval, nx.dsc, &D->gersh[l + 1]);
				rc = rc < 0 ? rc : (l == 0
					? BFMG_LAUNCH((k_mg_pscale<2, float>), (w.L.n_p + kBlock - 1) / kBlock, kBlock, 0, w.L, nx.dsc, (float*) w.pval)
					: BFMG_LAUNCH((k_mg_pscale<3, double>), (w.L.n_p + kBlock - 1) / kBlock, kBlock, 0, w.L, nx.dsc, (double*) w.pval));

				if (rc < 0) {
					return -1;
				}
			}
		}

		/* dense inverse of the last level (in place, coarse.cuh; replicated: every rank inverts it) */

		CoarseWork CW = {};

		CW.C.nc = nc;
		CW.C.half_bw = M->half_bw;
		CW.E = E;
		CW.P = Pblk;
		CW.bad = bad;

		if (coarse_invert(CW, S, false, 0, nc / kGjBlock) < 0) {
			return -1;
		}

		/* damping factors from the row-sum bounds - on several GPUs from the largest bound over the ranks, which also
		 * agree on whether a coarse operator came out unusable */

		int const stride = BFMG_MG_MAX_LEVELS + 1;

		if (dist && (
			BFMG_LAUNCH(k_mg_bounds, 1, kWarp, 0, (MgDev const*) D, (int32_t const*) bad, bounds + (size_t) world * stride) < 0 ||
			bfmg_dist_allgather_f64(bounds + (size_t) world * stride, bounds, stride) < 0
		)) {
			return -1;
		}

		if (BFMG_LAUNCH(k_mg_omega, 1, kWarp, 0, n_levels, omega_factor, D, (double const*) bounds, dist ? world : 0, stride) < 0) {
			return -1;
		}

		int32_t flags[2] = {0, 0};
		unsigned long long bound0 = 0;

		if (
			BFMG_CHECK(cudaMemcpyAsync(&flags[0], bad, sizeof(int32_t), cudaMemcpyDeviceToHost, bfmg_stream())) < 0 ||
			BFMG_CHECK(cudaMemcpyAsync(&flags[1], &D->bad, sizeof(int32_t), cudaMemcpyDeviceToHost, bfmg_stream())) < 0 ||
			BFMG_CHECK(cudaMemcpyAsync(&bound0, &D->gersh[0], sizeof bound0, cudaMemcpyDeviceToHost, bfmg_stream())) < 0 ||
			BFMG_CHECK(cudaStreamSynchronize(bfmg_stream())) < 0
		) {
			return -1;
		}

		memcpy(&lambda0, &bound0, sizeof lambda0);

		if (!(lambda0 >= 1) || !(lambda0 < 64)) {
			lambda0 = 4;
		}

		/* D->bad carries every rank's verdict after k_mg_omega; the dense inverse is replicated, its verdict the same everywhere */
		*usable = flags[0] == 0 && flags[1] == 0;
		return 0;
	}

	/* one cycle on the sparse level l >= 1 for the right-hand side W[l].g; returns the vector holding the result */
	double* cycle(int l, Scalars* S, bool obey) {
		MgLevelWork& w = W[l];
		bool const dense_next = l + 1 == n_levels - 1;
		double const* const om = &D->omega[l];

		if (!halo(l, w.g, S, obey) || BFMG_LAUNCH((k_blk_spmv<kMgPre, float>), w.grid_rows, kBlock, 0, w.L, (float const*) w.fval, w.g, w.g, w.vt, om, S, obey) < 0) {
			return nullptr;
		}

		for (int visit = 0; visit < w.gamma; visit++) {
			if (visit > 0 && (!halo(l, w.va, S, obey) || BFMG_LAUNCH((k_blk_spmv<kMgResid, float>), w.grid_rows, kBlock, 0, w.L, (float const*) w.fval, w.va, w.g, w.vt, om, S, obey) < 0)) {
				return nullptr;
			}

			if (!restrict_down(l, w.vt, S, obey)) {
				return nullptr;
			}

			double const* mu;

			if (dense_next) {
				if (BFMG_LAUNCH(k_dense_apply, (nc + kCoarseRows - 1) / kCoarseRows, kBlock, 0, nc, E, dg, dmu, S, obey) < 0) {
					return nullptr;
				}

				mu = dmu;
			}

			else {
				mu = cycle(l + 1, S, obey);

				if (mu == nullptr) {
					return nullptr;
				}
			}

			int const rc = visit == 0
				? BFMG_LAUNCH((k_mg_prolong<3, double, false>), w.grid_nodes, kBlock, 0, w.L, 0, w.L.row_hi, (double const*) w.pval, mu, w.g, w.va, om, S, obey)
				: BFMG_LAUNCH((k_mg_prolong<3, double, true>), w.grid_nodes, kBlock, 0, w.L, 0, w.L.row_hi, (double const*) w.pval, mu, w.g, w.va, om, S, obey);

			if (rc < 0) {
				return nullptr;
			}
		}

		if (!halo(l, w.va, S, obey) || BFMG_LAUNCH((k_blk_spmv<kMgPost, float>), w.grid_rows, kBlock, 0, w.L, (float const*) w.fval, w.va, w.g, w.vb, om, S, obey) < 0) {
			return nullptr;
		}

		return w.vb;
	}

	/* z = M^-1 r on level 0: t (work) and z are level-0 vectors; the result lands in `out` (may alias t) together
	 * with r.z -> beta, rho in S.  FIRST: initial residual (beta = 0, runs regardless of S->done). */
	template <bool FIRST>
	bool apply(bfmg_pattern_t const* pat, double2 const* stop, double2 const* sbot, double2* r, double2* t, double2* z, double2* out, double* partials, int spmv_grid, int vec_grid, Scalars* S) {
		bool const obey = !FIRST;
		bool const dense_next = n_levels == 2;
		double const* const om = &D->omega[0];
		int const lo = pat->row_lo;
		int const n_own = pat->row_hi - pat->row_lo;

		if (
			!halo(0, (double*) r, S, obey) ||
			BFMG_LAUNCH((k_spmv_mg<kMgPre, FIRST>), spmv_grid, kBlock, 0, *pat, (float2 const*) ftop, (float2 const*) fbot, (double2 const*) r, (double2 const*) r, t, om, partials, S) < 0 ||
			!restrict_down(0, (double const*) t, S, obey)
		) {
			return false;
		}

		double const* mu;

		if (dense_next) {
			if (BFMG_LAUNCH(k_dense_apply, (nc + kCoarseRows - 1) / kCoarseRows, kBlock, 0, nc, E, dg, dmu, S, obey) < 0) {
				return false;
			}

			mu = dmu;
		}

		else {
			mu = cycle(1, S, obey);

			if (mu == nullptr) {
				return false;
			}
		}

		if (
			BFMG_LAUNCH((k_mg_prolong<2, float, false>), vec_grid, kBlock, 0, W[0].L, lo, n_own, (float const*) W[0].pval, mu, (double const*) r, (double*) z, om, S, obey) < 0 ||
			!halo(0, (double*) z, S, obey) ||
			BFMG_LAUNCH((k_spmv_mg<kMgPost, FIRST>), spmv_grid, kBlock, 0, *pat, (float2 const*) ftop, (float2 const*) fbot, (double2 const*) z, (double2 const*) r, out, om, partials, S) < 0
		) {
			return false;
		}

		/* several GPUs: r.z, r.r and x.x of all ranks, folded in rank order by one thread per rank */
		return !dist || (FIRST ? BFMG_LAUNCH(k_share<kShareMgFirst>, 1, 1, 0, S, 0) : BFMG_LAUNCH(k_share<kShareMg>, 1, 1, 0, S, 0)) == 0;
	}
};

} // namespace
