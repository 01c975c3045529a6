/*
 * csr.c - BFM_MATRIX_KIND_CSR: the device-resident sparse matrix behind bfm_matrix_t.
 *
 * Values live in GPU memory in the plan's SELL-32 node-block layout (internal.h).  The host-facing
 * accessors (get/set/add/bandwidth/RCM) work on a lazily refreshed host mirror; bfm_matrix_solve
 * runs the FP64 PCG kernels (solver.cu).  A renumbering applied through bfm_perm_perm_matrix is kept
 * as a logical permutation: the 2x2 node blocks the kernels rely on would not survive a DOF-level
 * reordering, and CG does not need the matrix physically permuted.
 */
#include "internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline bfmi_csr_t* impl_of(bfm_matrix_t const* matrix) {
	return (bfmi_csr_t*) matrix->csr.impl;
}

/* position of scalar entry (i, j) inside the value array, or -1 (ORIGINAL numbering) */
static inline int64_t value_index(bfmi_csr_t const* csr, size_t i, size_t j) {
	int64_t const slot = bfmi_plan_find(csr->plan, (int32_t) (i / 2), (int32_t) (j / 2));

	if (slot < 0) {
		return -1;
	}

	return (int64_t) (i % 2) * 2 * csr->plan->n_slots + 2 * slot + (int64_t) (j % 2);
}

int bfmi_csr_wrap(bfm_matrix_t* matrix, bfm_state_t* state, bfmi_plan_t* plan, double* d_val) {
	bfmi_csr_t* const csr = calloc(1, sizeof *csr);

	if (csr == NULL) {
		return -1;
	}

	csr->state = state;
	csr->plan = plan;
	csr->d_val = d_val;
	csr->d_valid = d_val != NULL;

	bfmi_plan_retain(plan);

	matrix->state = state;
	matrix->kind = BFM_MATRIX_KIND_CSR;
	matrix->major = BFM_MATRIX_MAJOR_ROW;
	matrix->m = 2 * (size_t) plan->nb;
	matrix->csr.impl = csr;
	matrix->csr.reserved = NULL;

	return 0;
}

int bfmi_csr_destroy(bfm_matrix_t* matrix) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (csr == NULL) {
		return 0;
	}

	bfmg_free(csr->d_val);
	free(csr->h_val);
	free(csr->perm);
	free(csr->inv_perm);
	bfmi_plan_release(csr->plan);
	free(csr);

	matrix->csr.impl = NULL;
	return 0;
}

/* make the device copy current (matrices created from host data are uploaded on first use) */
static int ensure_device(bfmi_csr_t* csr) {
	if (csr->d_valid) {
		return 0;
	}

	size_t const bytes = (size_t) csr->plan->n_slots * 4 * sizeof(double);

	if (!csr->h_valid || bfmi_plan_upload(csr->state, csr->plan) < 0) {
		return -1;
	}

	if (csr->d_val == NULL && bfmg_alloc((void**) &csr->d_val, bytes) < 0) {
		return BFMI_FAIL(csr->state, "%s", bfmg_last_error());
	}

	if (bfmg_upload(csr->d_val, csr->h_val, bytes) < 0 || bfmg_sync() < 0) {
		return BFMI_FAIL(csr->state, "%s", bfmg_last_error());
	}

	csr->d_valid = true;
	return 0;
}

int bfmi_csr_mirror(bfmi_csr_t* csr) {
	if (csr->h_valid) {
		return 0;
	}

	if (!csr->d_valid) {
		return -1;
	}

	size_t const bytes = (size_t) csr->plan->n_slots * 4 * sizeof(double);

	if (csr->h_val == NULL) {
		csr->h_val = malloc(bytes ? bytes : 8);

		if (csr->h_val == NULL) {
			return -1;
		}
	}

	if (bfmg_download(csr->h_val, csr->d_val, bytes) < 0) {
		return BFMI_FAIL(csr->state, "reading the matrix back failed: %s", bfmg_last_error());
	}

	csr->h_valid = true;
	return 0;
}

double bfmi_csr_get(bfm_matrix_t* matrix, size_t i, size_t j) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (i >= matrix->m || j >= matrix->m || bfmi_csr_mirror(csr) < 0) {
		return BFM_NAN;
	}

	if (csr->inv_perm != NULL) { /* A'[perm[i]][perm[j]] = A[i][j] */
		i = csr->inv_perm[i];
		j = csr->inv_perm[j];
	}

	int64_t const at = value_index(csr, i, j);
	return at < 0 ? 0 : csr->h_val[at];
}

int bfmi_csr_put(bfm_matrix_t* matrix, size_t i, size_t j, double val, bool add) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (i >= matrix->m || j >= matrix->m || bfmi_csr_mirror(csr) < 0) {
		return -1;
	}

	if (csr->inv_perm != NULL) {
		i = csr->inv_perm[i];
		j = csr->inv_perm[j];
	}

	int64_t const at = value_index(csr, i, j);

	if (at < 0) { /* outside the sparsity pattern: same rule as a write outside a band */
		return fabs(val) < BFM_PIVOT_EPS ? 0 : -1;
	}

	csr->h_val[at] = add ? csr->h_val[at] + val : val;

	if (csr->d_valid && (bfmg_upload(csr->d_val + at, &csr->h_val[at], sizeof(double)) < 0 || bfmg_sync() < 0)) {
		return -1;
	}

	return 0;
}

int bfmi_csr_copy(bfm_matrix_t* dst, bfm_matrix_t* src) {
	bfmi_csr_t* const d = impl_of(dst);
	bfmi_csr_t* const s = impl_of(src);

	if (d->plan != s->plan || d->perm != NULL || s->perm != NULL || ensure_device(s) < 0 || ensure_device(d) < 0) {
		return -1;
	}

	d->h_valid = false;
	return bfmg_copy(d->d_val, s->d_val, (size_t) s->plan->n_slots * 4 * sizeof(double));
}

size_t bfmi_csr_bandwidth(bfm_matrix_t* matrix) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (bfmi_csr_mirror(csr) < 0) {
		return (size_t) -1;
	}

	bfmi_plan_t const* const plan = csr->plan;
	size_t k = 0;

	for (int32_t a = 0; a < plan->nb; a++) {
		for (int32_t t = 0; t < plan->row_len[a]; t++) {
			int64_t const slot = (int64_t) plan->slice_off[a / 32] + (int64_t) t * 32 + a % 32;
			size_t const b = (size_t) plan->scol[slot];

			for (size_t r = 0; r < 2; r++) {
				for (size_t c = 0; c < 2; c++) {
					if (csr->h_val[r * 2 * plan->n_slots + 2 * slot + c] == 0) {
						continue;
					}

					size_t i = 2 * (size_t) a + r;
					size_t j = 2 * b + c;

					if (csr->perm != NULL) {
						i = csr->perm[i];
						j = csr->perm[j];
					}

					size_t const d = i > j ? i - j : j - i;
					k = d > k ? d : k;
				}
			}
		}
	}

	return k;
}

/* ---- RCM front end (perm.c holds the traversal) ------------------------------------------------ */

static size_t csr_row_nnz(size_t i, size_t* out, void* ctx) {
	bfmi_csr_t const* const csr = ctx;
	bfmi_plan_t const* const plan = csr->plan;

	int32_t const a = (int32_t) (i / 2);
	size_t const r = i % 2;
	size_t cnt = 0;

	for (int32_t t = 0; t < plan->row_len[a]; t++) {
		int64_t const slot = (int64_t) plan->slice_off[a / 32] + (int64_t) t * 32 + a % 32;
		double const* const v = &csr->h_val[r * 2 * plan->n_slots + 2 * slot];

		if (v[0] != 0) {
			out[cnt++] = 2 * (size_t) plan->scol[slot];
		}

		if (v[1] != 0) {
			out[cnt++] = 2 * (size_t) plan->scol[slot] + 1;
		}
	}

	return cnt;
}

int bfmi_csr_rcm(bfm_perm_t* perm, bfm_matrix_t* matrix) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (csr->perm != NULL) {
		return -1; /* renumbering an already renumbered matrix is not supported */
	}

	if (bfmi_csr_mirror(csr) < 0) {
		return -1;
	}

	size_t max_row = 0;

	for (int32_t a = 0; a < csr->plan->nb; a++) {
		size_t const c = 2 * (size_t) csr->plan->row_len[a];
		max_row = c > max_row ? c : max_row;
	}

	return bfmi_rcm(perm, matrix->m, max_row, csr_row_nnz, csr);
}

int bfmi_csr_set_perm(bfm_matrix_t* matrix, size_t const* perm, size_t n) {
	bfmi_csr_t* const csr = impl_of(matrix);

	if (n != matrix->m || csr->perm != NULL) {
		return -1;
	}

	/* validated and built aside: a failure leaves the matrix exactly as it was */

	size_t* const fwd = malloc((n + 1) * sizeof *fwd);
	size_t* const inv = malloc((n + 1) * sizeof *inv);

	if (fwd == NULL || inv == NULL) {
		free(fwd);
		free(inv);
		return -1;
	}

	for (size_t i = 0; i < n; i++) {
		inv[i] = SIZE_MAX;
	}

	for (size_t i = 0; i < n; i++) {
		if (perm[i] >= n || inv[perm[i]] != SIZE_MAX) { /* out of range, or not a permutation */
			free(fwd);
			free(inv);
			return -1;
		}

		fwd[i] = perm[i];
		inv[perm[i]] = i;
	}

	csr->perm = fwd;
	csr->inv_perm = inv;

	return 0;
}

/* ---- solve ------------------------------------------------------------------------------------- */

void bfmi_pcg_options(size_t n, bfmg_pcg_opts_t* opts) {
	char const* env;

	opts->tol = 1e-12; /* BASELINE.json north_star: CG stopped at a relative residual <= 1e-12 */
	opts->max_iter = (int32_t) (100 * sqrt((double) n) + 10000);
	opts->chunk = 0; /* the solver's default: 64, or 8 with the multilevel preconditioner */
	opts->verify = 1;
	opts->true_tol = 1e-10; /* normwise backward error, see solver.cu */
	opts->max_restarts = 0;

	if ((env = getenv("BFM_CG_TOL")) != NULL && atof(env) > 0) {
		opts->tol = atof(env);
	}

	if ((env = getenv("BFM_CG_MAXIT")) != NULL && atoi(env) > 0) {
		opts->max_iter = atoi(env);
	}

	if ((env = getenv("BFM_CG_CHUNK")) != NULL && atoi(env) > 0) {
		opts->chunk = atoi(env);
	}

	if ((env = getenv("BFM_CG_VERIFY")) != NULL) {
		opts->verify = atoi(env);
	}

	if ((env = getenv("BFM_CG_TRUE_TOL")) != NULL && atof(env) > 0) {
		opts->true_tol = atof(env);
	}

	if ((env = getenv("BFM_CG_RESTARTS")) != NULL) {
		opts->max_restarts = atoi(env);
	}
}

int bfmi_csr_solve(bfm_matrix_t* matrix, bfm_vec_t* y) {
	bfmi_csr_t* const csr = impl_of(matrix);
	bfm_state_t* const state = csr->state;
	size_t const n = matrix->m;

	if (bfmi_plan_upload(state, csr->plan) < 0 || ensure_device(csr) < 0) {
		return -1;
	}

	double* d_b = NULL;
	double* d_x = NULL;
	double* staged = NULL;
	int rv = -1;

	if (bfmg_alloc((void**) &d_b, n * sizeof *d_b) < 0 || bfmg_alloc((void**) &d_x, n * sizeof *d_x) < 0) {
		BFMI_FAIL(state, "%s", bfmg_last_error());
		goto done;
	}

	/* the caller's vector is in the renumbered order when a permutation is active:
	 * b'[perm[i]] = b[i]  (reference perm.c:97-100), so b[i] = b'[perm[i]] */

	double const* rhs = y->data;

	if (csr->perm != NULL) {
		staged = malloc(n * sizeof *staged);

		if (staged == NULL) {
			goto done;
		}

		for (size_t i = 0; i < n; i++) {
			staged[i] = y->data[csr->perm[i]];
		}

		rhs = staged;
	}

	bfmg_pcg_opts_t opts;
	bfmg_pcg_result_t res = {0};

	bfmi_pcg_options(n, &opts);

	if (bfmg_upload(d_b, rhs, n * sizeof *d_b) < 0) {
		BFMI_FAIL(state, "%s", bfmg_last_error());
		goto done;
	}

	/* small systems: the whole PCG inside one CTA, as bfm_sim_run does (batch.cu); the coarse level needs
	 * node coordinates, which a bare matrix does not carry, so large ones get the diagonal preconditioner */

	char const* const one_cta = getenv("BFM_ONE_CTA");
	int solved;

	if (csr->plan->nb <= bfmg_batch_max_rows() && (one_cta == NULL || atoi(one_cta) != 0)) {
		bfmg_batch_range_t const range = {0, csr->plan->nb};
		bfmg_batch_status_t st;
		size_t const before = bfmg_launch_count();

		solved = bfmg_pcg_batch(&csr->plan->dev, csr->d_val, d_b, d_x, &opts, 1, &range, &st, &res.ms);

		res.iterations = st.iterations;
		res.rel_residual = st.rel_residual;
		res.true_rel_residual = st.true_rel_residual;
		res.backward_error = st.backward_error;
		res.converged = st.converged == 1 && st.backward_error > opts.true_tol ? 0 : st.converged;
		res.launches = bfmg_launch_count() - before;
	}

	else {
		solved = bfmg_pcg(&csr->plan->dev, csr->d_val, d_b, d_x, &opts, &res, NULL, NULL, NULL);
	}

	if (solved < 0) {
		BFMI_FAIL(state, "PCG failed: %s", bfmg_last_error());
		goto done;
	}

	double* const out = staged != NULL ? staged : y->data;

	if (bfmg_download(out, d_x, n * sizeof *d_x) < 0) {
		BFMI_FAIL(state, "%s", bfmg_last_error());
		goto done;
	}

	if (csr->perm != NULL) {
		for (size_t i = 0; i < n; i++) {
			y->data[csr->perm[i]] = staged[i];
		}
	}

	memset(&csr->stats, 0, sizeof csr->stats);

	csr->stats.n_dofs = n;
	csr->stats.n_blocks = (size_t) csr->plan->n_blocks;
	csr->stats.n_slots = (size_t) csr->plan->n_slots;
	csr->stats.cg_iterations = res.iterations;
	csr->stats.cg_restarts = res.restarts;
	csr->stats.cg_converged = res.converged;
	csr->stats.cg_rel_residual = res.rel_residual;
	csr->stats.cg_true_rel_residual = res.true_rel_residual;
	csr->stats.cg_backward_error = res.backward_error;
	csr->stats.ms_solve = res.ms;
	csr->stats.kernel_launches = res.launches;
	csr->stats.h2d_bytes = n * sizeof(double);
	csr->stats.d2h_bytes = n * sizeof(double);

	bfmx_publish_stats(&csr->stats);

	if (res.converged != 1) {
		BFMI_FAIL(state, "PCG stopped after %d iterations without converging (relative residual %.3e, true %.3e)", res.iterations, res.rel_residual, res.true_rel_residual);
		goto done;
	}

	rv = 0;

done:

	free(staged);
	bfmg_free(d_b);
	bfmg_free(d_x);

	return rv;
}

/* ---- user-supplied sparse matrices -------------------------------------------------------------- */

int bfmx_matrix_csr_create(bfm_matrix_t* matrix, bfm_state_t* state, size_t n, size_t const* rowptr, size_t const* col, double const* val) {
	bfmi_plan_t* const plan = bfmi_plan_from_csr(state, n, rowptr, col);

	if (plan == NULL) {
		return BFMI_FAIL(state, "bfmx_matrix_csr_create: n must be even and the pattern valid");
	}

	size_t const count = (size_t) plan->n_slots * 4;
	double* const h_val = calloc(count ? count : 1, sizeof *h_val);

	if (h_val == NULL) {
		bfmi_plan_release(plan);
		return -1;
	}

	for (size_t i = 0; i < n; i++) {
		for (size_t t = rowptr[i]; t < rowptr[i + 1]; t++) {
			int64_t const slot = bfmi_plan_find(plan, (int32_t) (i / 2), (int32_t) (col[t] / 2));
			h_val[(i % 2) * 2 * (size_t) plan->n_slots + 2 * (size_t) slot + col[t] % 2] += val[t];
		}
	}

	/* uploaded lazily by the first solve, so that host-only use needs no device */

	if (bfmi_csr_wrap(matrix, state, plan, NULL) < 0) {
		free(h_val);
		bfmi_plan_release(plan);

		return -1;
	}

	bfmi_csr_t* const csr = impl_of(matrix);

	csr->h_val = h_val;
	csr->h_valid = true;

	bfmi_plan_release(plan); /* the matrix holds its own reference */
	return 0;
}

/* scalar CSR view (structural pattern, ascending columns, ORIGINAL numbering); pass NULL arrays to
 * query the entry count only */
int bfmx_matrix_csr_export(bfm_matrix_t* matrix, size_t* nnz, size_t* rowptr, size_t* col, double* val) {
	if (matrix->kind != BFM_MATRIX_KIND_CSR) {
		return -1;
	}

	bfmi_csr_t* const csr = impl_of(matrix);
	bfmi_plan_t const* const plan = csr->plan;

	*nnz = 4 * (size_t) plan->n_blocks;

	if (rowptr == NULL || col == NULL || val == NULL) {
		return 0;
	}

	if (bfmi_csr_mirror(csr) < 0) {
		return -1;
	}

	size_t at = 0;

	for (int32_t a = 0; a < plan->nb; a++) {
		for (size_t r = 0; r < 2; r++) {
			rowptr[2 * (size_t) a + r] = at;

			for (int32_t t = 0; t < plan->row_len[a]; t++) {
				int64_t const slot = (int64_t) plan->slice_off[a / 32] + (int64_t) t * 32 + a % 32;

				for (size_t c = 0; c < 2; c++) {
					col[at] = 2 * (size_t) plan->scol[slot] + c;
					val[at++] = csr->h_val[r * 2 * (size_t) plan->n_slots + 2 * (size_t) slot + c];
				}
			}
		}
	}

	rowptr[matrix->m] = at;
	return 0;
}
