/*
 * assembly.cu - element stiffness/load assembly and boundary conditions on sm_100a.
 *
 * COMPILED WITH -fmad=false.  The reference accumulates every matrix entry as a chain of IEEE double
 * additions - elements ascending, integration points ascending, one add per point (system.c:136,
 * :207-223, :460) - and bfm's "sparsity pattern" is the set of entries that are NOT exact zeros after
 * that chain (perm.c:140-144, :262).  Whether a cancellation such as t+t+t-t-t-t lands on exactly 0
 * depends on the order and on the absence of fused multiply-adds, so this kernel reproduces both:
 * each lane owns one 2x2 node block and walks its precomputed contributor list (the deterministic
 * element-to-nonzero map) sequentially, in reference order.  No atomics, no shuffle trees - the
 * "segmented reduction" is one ordered segment per lane, which is what bit-for-bit parity requires.
 *
 * Layout: SELL-32 node blocks (internal.h).  A warp takes one 32-row slice, lane = block row, and
 * steps through the slice's slots, so the contributor-pointer reads and the value writes are
 * contiguous 32-wide segments; coordinates and connectivity are gathered through L1/L2 (each is
 * reused by the ~7-9 blocks of a row and by neighbouring rows).
 *
 * Roofline: HBM.  Algorithmic bytes per P1 element: 12 (connectivity) + 8 (coords, amortised) +
 * 36 (contributor map) + 4*4*... see DESIGN.md section "Kernels".
 */
#include "gpu_internal.cuh"

namespace {

template <int KIND>
__device__ __forceinline__ double pick(double const (&v)[KIND], int idx) {
	double out = v[0];

#pragma unroll
	for (int j = 1; j < KIND; j++) {
		out = idx == j ? v[j] : out;
	}

	return out;
}

/* Geometry of one element at one integration point: |det J| and the physical gradients of all KIND
 * shape functions (reference system.c:164-186), plus the radius for the axisymmetric form (:302). */
template <int KIND>
struct Geom {
	double det;
	double r;
	double dpx[KIND];
	double dpy[KIND];
};

template <int KIND>
__device__ __forceinline__ void geometry(bfmg_asm_tables_t const& T, int g, double const (&x)[KIND], double const (&y)[KIND], Geom<KIND>& G) {
	double dx_dxsi = 0, dx_deta = 0, dy_dxsi = 0, dy_deta = 0, r = 0;

#pragma unroll
	for (int j = 0; j < KIND; j++) {
		dx_dxsi += x[j] * T.dxsi[g][j];
		dx_deta += x[j] * T.deta[g][j];
		dy_dxsi += y[j] * T.dxsi[g][j];
		dy_deta += y[j] * T.deta[g][j];
		r += x[j] * T.phi[g][j];
	}

	G.det = fabs(dx_dxsi * dy_deta - dx_deta * dy_dxsi);
	G.r = r;

#pragma unroll
	for (int j = 0; j < KIND; j++) {
		G.dpx[j] = (T.dxsi[g][j] * dy_deta - T.deta[g][j] * dy_dxsi) / G.det;
		G.dpy[j] = (T.deta[g][j] * dx_dxsi - T.dxsi[g][j] * dx_deta) / G.det;
	}
}

/* BATCH: the pattern holds many independent systems (batch.cu); slice s belongs to system
 * slice_tab[s], whose tables (material, forces, rule) are tabs[slice_tab[s]] in device memory */
template <int KIND, bool AXI, bool BATCH>
__global__ void __launch_bounds__(kBlock) k_assemble(
	const __grid_constant__ bfmg_asm_tables_t T0, const __grid_constant__ bfmg_pattern_t P,
	double2 const* __restrict__ coords, double2 const* __restrict__ nforce,
	double2* __restrict__ vtop, double2* __restrict__ vbot, double2* __restrict__ bvec,
	bfmg_asm_tables_t const* __restrict__ tabs, int32_t const* __restrict__ slice_tab
) {
	pdl_sync();

	int const lane = threadIdx.x & (kWarp - 1);
	int const warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
	int const n_warps = gridDim.x * blockDim.x / kWarp;

	for (int slice = warp; slice < P.n_slices; slice += n_warps) {
		int const row = slice * kWarp + lane;
		int const diag = row < P.nb ? P.diag_pos[row] : -1;
		int const end = P.slice_off[slice + 1];

		bfmg_asm_tables_t const& T = BATCH ? tabs[slice_tab[slice]] : T0;

		for (int slot = P.slice_off[slice] + lane; slot < end; slot += kWarp) {
			int const c_end = P.ctr_ptr[slot + 1];
			bool const on_diag = slot == diag;

			double a11 = 0, a12 = 0, a21 = 0, a22 = 0; /* the 2x2 block */
			double b0 = 0, b1 = 0;                     /* load of this node (diagonal lanes only) */

			for (int c = P.ctr_ptr[slot]; c < c_end;) {
				uint32_t const first = P.ctr[c];
				int const elem = first >> 4;

				/* contributions of one element to this block: normally exactly one (j, k) pair;
				 * several only if the element repeats a node */

				int c_next = c + 1;

				while (c_next < c_end && (int) (P.ctr[c_next] >> 4) == elem) {
					c_next++;
				}

				int node[KIND];
				double x[KIND], y[KIND];

#pragma unroll
				for (int j = 0; j < KIND; j++) { /* get_elem, system.c:93-107 */
					node[j] = P.elems[elem * KIND + j];
					double2 const xy = coords[node[j]];
					x[j] = xy.x;
					y[j] = xy.y;
				}

				Geom<KIND> G;

				for (int g = 0; g < T.n_points; g++) { /* system.c:136 */
					if (g == 0 || !T.grad_const) {
						geometry<KIND>(T, g, x, y, G);
					}

					if (AXI && T.grad_const) { /* the radius still moves with the point */
						double r = 0;

#pragma unroll
						for (int j = 0; j < KIND; j++) {
							r += x[j] * T.phi[g][j];
						}

						G.r = r;
					}

					double const dw = G.det * T.weight[g];

					for (int t = c; t < c_next; t++) {
						uint32_t const packed = P.ctr[t];
						int const j = (packed >> 2) & 3;
						int const k = packed & 3;

						double const phi_j = T.phi[g][j];

						if (on_diag && j == k) { /* load vector, system.c:190-203 / :319-332 */
							for (int f = 0; f < T.n_forces; f++) {
								double2 F;

								if (T.forces_per_node) {
									F = nforce[(size_t) f * P.nb + row];
								}

								else {
									F = make_double2(T.const_force[f][0], T.const_force[f][1]);
								}

								if (AXI) {
									b0 += dw * F.x * T.rho * phi_j * G.r;
									b1 += dw * F.y * T.rho * phi_j * G.r;
								}

								else {
									b0 += dw * F.x * T.rho * phi_j;
									b1 += dw * F.y * T.rho * phi_j;
								}
							}
						}

						double const dxj = pick<KIND>(G.dpx, j), dyj = pick<KIND>(G.dpy, j);
						double const dxk = pick<KIND>(G.dpx, k), dyk = pick<KIND>(G.dpy, k);

						double f11, f12, f21, f22;

						if (AXI) { /* system.c:342-345 */
							double const r = G.r;
							double const phi_k = T.phi[g][k];

							f11 = T.a * dxj * dxk * r + T.c * dyj * dyk * r + phi_j * (T.b * dxk + T.a * phi_k / r) + dxj * T.b * phi_k;
							f12 = T.b * dxj * dyk * r + T.c * dyj * dxk * r + phi_j * T.b * dyk;
							f21 = T.b * dyj * dxk * r + T.c * dxj * dyk * r + dyj * T.b * phi_k;
							f22 = T.a * dyj * dyk * r + T.c * dxj * dxk;
						}

						else { /* system.c:213-216 */
							f11 = T.a * dxj * dxk + T.c * dyj * dyk;
							f12 = T.b * dxj * dyk + T.c * dyj * dxk;
							f21 = T.b * dyj * dxk + T.c * dxj * dyk;
							f22 = T.a * dyj * dyk + T.c * dxj * dxk;
						}

						a11 += dw * f11; /* system.c:218-221 */
						a12 += dw * f12;
						a21 += dw * f21;
						a22 += dw * f22;
					}
				}

				c = c_next;
			}

			vtop[slot] = make_double2(a11, a12);
			vbot[slot] = make_double2(a21, a22);

			if (on_diag) {
				bvec[row] = make_double2(b0, b1);
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * Dirichlet-type conditions.  The reference applies apply_constraint (system.c:358-374) once per
 * constrained DOF d, in ascending DOF order within a condition:
 *     for every row i:  b[i] -= value_d * A[i][d];  A[i][d] = 0
 *     row d = 0;  A[d][d] = 1;  b[d] = value_d
 * Per row that is an ordered walk over the row's constrained columns - columns are stored
 * ascending, so one lane per block row replays it exactly - followed, for a row that is itself
 * constrained, by the identity row.  Only rows adjacent to a constrained node are launched.
 * ---------------------------------------------------------------------------------------------- */

__global__ void k_bc_mark(int32_t* __restrict__ stamp, double* __restrict__ cval, int32_t epoch, int32_t const* __restrict__ dofs, double const* __restrict__ vals, int n) {
	pdl_sync();

	int const i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < n) {
		stamp[dofs[i]] = epoch;
		cval[dofs[i]] = vals[i];
	}
}

__global__ void __launch_bounds__(kBlock) k_bc_dirichlet(
	const __grid_constant__ bfmg_pattern_t P, double2* __restrict__ vtop, double2* __restrict__ vbot, double2* __restrict__ bvec,
	int32_t const* __restrict__ stamp, double const* __restrict__ cval, int32_t epoch, int32_t const* __restrict__ rows, int n_rows
) {
	pdl_sync();

	int const i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i >= n_rows) {
		return;
	}

	int const row = rows[i];
	int const base = P.slice_off[row / kWarp] + row % kWarp;
	int const len = P.row_len[row];

	double2 b = bvec[row];

	bool const fix0 = stamp[2 * row + 0] == epoch;
	bool const fix1 = stamp[2 * row + 1] == epoch;

	for (int t = 0; t < len; t++) {
		int const slot = base + t * kWarp;
		int const col = P.scol[slot];

		bool const c0 = stamp[2 * col + 0] == epoch;
		bool const c1 = stamp[2 * col + 1] == epoch;

		if (!c0 && !c1 && !fix0 && !fix1) {
			continue;
		}

		double2 top = vtop[slot];
		double2 bot = vbot[slot];

		if (c0) {
			double const v = cval[2 * col + 0];

			b.x -= v * top.x;
			b.y -= v * bot.x;
			top.x = 0;
			bot.x = 0;
		}

		if (c1) {
			double const v = cval[2 * col + 1];

			b.x -= v * top.y;
			b.y -= v * bot.y;
			top.y = 0;
			bot.y = 0;
		}

		if (fix0) {
			top = make_double2(col == row ? 1 : 0, 0);
		}

		if (fix1) {
			bot = make_double2(0, col == row ? 1 : 0);
		}

		vtop[slot] = top;
		vbot[slot] = bot;
	}

	if (fix0) {
		b.x = cval[2 * row + 0];
	}

	if (fix1) {
		b.y = cval[2 * row + 1];
	}

	bvec[row] = b;
}

__global__ void k_bc_add(double* __restrict__ bvec, int32_t const* __restrict__ group_dof, int32_t const* __restrict__ group_ptr, double const* __restrict__ add, int n_groups) {
	pdl_sync();

	int const i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i >= n_groups) {
		return;
	}

	double acc = bvec[group_dof[i]];

	for (int t = group_ptr[i]; t < group_ptr[i + 1]; t++) {
		acc += add[t];
	}

	bvec[group_dof[i]] = acc;
}

} // namespace

extern "C" {

int bfmg_assemble(bfmg_pattern_t const* pat, bfmg_asm_tables_t const* tab, double const* d_coords, double const* d_nforce, double* d_val, double* d_b, bfmg_asm_tables_t const* d_tabs, int32_t const* d_slice_tab) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (pat->n_slices == 0) {
		return 0;
	}

	double2* const vtop = (double2*) d_val;
	double2* const vbot = vtop + pat->n_slots;

	int const blocks_needed = (pat->n_slices + kWarpsPerBlock - 1) / kWarpsPerBlock;

#define ASM_LAUNCH(KIND, AXI, BATCH) BFMG_LAUNCH((k_assemble<KIND, AXI, BATCH>), bfmg_grid(blocks_needed, bfmg_resident_ctas(k_assemble<KIND, AXI, BATCH>)), kBlock, 0, *tab, *pat, (double2 const*) d_coords, (double2 const*) d_nforce, vtop, vbot, (double2*) d_b, d_tabs, d_slice_tab)
#define ASM_PICK(KIND, AXI) (d_tabs != nullptr ? ASM_LAUNCH(KIND, AXI, true) : ASM_LAUNCH(KIND, AXI, false))

	/* tab->kind / tab->axisym are common to a batch (job.c checks) */

	if (tab->kind == 3) {
		return tab->axisym ? ASM_PICK(3, true) : ASM_PICK(3, false);
	}

	if (tab->kind == 4) {
		return tab->axisym ? ASM_PICK(4, true) : ASM_PICK(4, false);
	}

#undef ASM_PICK

#undef ASM_LAUNCH

	return -1;
}

int bfmg_bc_dirichlet(bfmg_pattern_t const* pat, double* d_val, double* d_b, int32_t* d_stamp, double* d_cval, int32_t epoch, int32_t const* d_dofs, double const* d_vals, int32_t n_dofs, int32_t const* d_rows, int32_t n_rows) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (n_dofs == 0) {
		return 0;
	}

	if (BFMG_LAUNCH(k_bc_mark, (n_dofs + kBlock - 1) / kBlock, kBlock, 0, d_stamp, d_cval, epoch, d_dofs, d_vals, n_dofs) < 0) {
		return -1;
	}

	double2* const vtop = (double2*) d_val;
	double2* const vbot = vtop + pat->n_slots;

	return BFMG_LAUNCH(k_bc_dirichlet, (n_rows + kBlock - 1) / kBlock, kBlock, 0, *pat, vtop, vbot, (double2*) d_b, d_stamp, d_cval, epoch, d_rows, n_rows);
}

int bfmg_bc_add(double* d_b, int32_t const* d_group_dof, int32_t const* d_group_ptr, double const* d_add, int32_t n_groups) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (n_groups == 0) {
		return 0;
	}

	return BFMG_LAUNCH(k_bc_add, (n_groups + kBlock - 1) / kBlock, kBlock, 0, d_b, d_group_dof, d_group_ptr, d_add, n_groups);
}

} // extern "C"
