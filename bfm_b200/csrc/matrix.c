/*
 * matrix.c - the public matrix API (bfm/matrix.h): host FULL and BAND kinds with the reference's
 * storage and arithmetic, and dispatch to the device-resident CSR kind (csr.c).
 *
 * FULL/BAND are SURVEY.md section 8(f) item 4: callers that drive bfm_system_create() +
 * bfm_matrix_add() by hand keep working.  The GPU hot path never touches them.
 * Replaces reference matrix.c:11-584 (the dead WITH_BLAS branches are not reproduced).
 */
#include "internal.h"

#include <string.h>

/* ---- storage ---------------------------------------------------------------------------------- */

static inline size_t full_index(bfm_matrix_t const* m, size_t i, size_t j) {
	return m->major == BFM_MATRIX_MAJOR_ROW ? i * m->m + j : i + j * m->m;
}

/* skewed band addressing of the reference: stride 2k inside an m * (2k + 1) buffer (matrix.c:208) */
static inline size_t band_index(bfm_matrix_t const* m, size_t i, size_t j) {
	return m->major == BFM_MATRIX_MAJOR_ROW ? j + i * 2 * m->band.k : i + j * 2 * m->band.k;
}

static inline bool in_band(bfm_matrix_t const* m, size_t i, size_t j) {
	size_t const d = i > j ? i - j : j - i;
	return d <= m->band.k;
}

static int create(bfm_matrix_t* matrix, bfm_state_t* state, bfm_matrix_kind_t kind, bfm_matrix_major_t major, size_t m, size_t count, double** data) {
	matrix->state = state;
	matrix->kind = kind;
	matrix->major = major;
	matrix->m = m;

	*data = state->alloc(count * sizeof **data);

	if (*data == NULL) {
		return -1;
	}

	memset(*data, 0, count * sizeof **data);
	return 0;
}

int bfm_matrix_full_create(bfm_matrix_t* matrix, bfm_state_t* state, bfm_matrix_major_t major, size_t m) {
	return create(matrix, state, BFM_MATRIX_KIND_FULL, major, m, m * m, &matrix->full.data);
}

int bfm_matrix_band_create(bfm_matrix_t* matrix, bfm_state_t* state, bfm_matrix_major_t major, size_t m, size_t k) {
	matrix->band.k = k;
	return create(matrix, state, BFM_MATRIX_KIND_BAND, major, m, m * (2 * k + 1), &matrix->band.data);
}

int bfm_matrix_destroy(bfm_matrix_t* matrix) {
	switch (matrix->kind) {
	case BFM_MATRIX_KIND_FULL:
		matrix->state->free(matrix->full.data);
		return 0;

	case BFM_MATRIX_KIND_BAND:
		matrix->state->free(matrix->band.data);
		return 0;

	case BFM_MATRIX_KIND_CSR:
		return bfmi_csr_destroy(matrix);
	}

	return -1;
}

/* ---- element access ---------------------------------------------------------------------------- */

double bfm_matrix_get(bfm_matrix_t* matrix, size_t i, size_t j) {
	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_get(matrix, i, j);
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL && matrix->kind != BFM_MATRIX_KIND_BAND) {
		return -1; /* matrix.c:464 */
	}

	if (i >= matrix->m || j >= matrix->m) {
		return BFM_NAN;
	}

	if (matrix->kind == BFM_MATRIX_KIND_FULL) {
		return matrix->full.data[full_index(matrix, i, j)];
	}

	return in_band(matrix, i, j) ? matrix->band.data[band_index(matrix, i, j)] : 0;
}

/* writes outside the band are accepted only for (near-)zeros (matrix.c:221-223, :239-241) */
static int put(bfm_matrix_t* matrix, size_t i, size_t j, double val, bool add) {
	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_put(matrix, i, j, val, add);
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL && matrix->kind != BFM_MATRIX_KIND_BAND) {
		return -1;
	}

	if (i >= matrix->m || j >= matrix->m) {
		return -1;
	}

	double* at;

	if (matrix->kind == BFM_MATRIX_KIND_FULL) {
		at = &matrix->full.data[full_index(matrix, i, j)];
	}

	else if (in_band(matrix, i, j)) {
		at = &matrix->band.data[band_index(matrix, i, j)];
	}

	else {
		return fabs(val) < BFM_PIVOT_EPS ? 0 : -1;
	}

	*at = add ? *at + val : val;
	return 0;
}

int bfm_matrix_set(bfm_matrix_t* matrix, size_t i, size_t j, double val) {
	return put(matrix, i, j, val, false);
}

int bfm_matrix_add(bfm_matrix_t* matrix, size_t i, size_t j, double val) {
	return put(matrix, i, j, val, true);
}

int bfm_matrix_copy(bfm_matrix_t* matrix, bfm_matrix_t* src) {
	if (matrix->m != src->m) {
		return -1;
	}

	if (matrix->kind == BFM_MATRIX_KIND_FULL && src->kind == BFM_MATRIX_KIND_FULL) {
		memcpy(matrix->full.data, src->full.data, src->m * src->m * sizeof *src->full.data);
		return 0;
	}

	if (matrix->kind == BFM_MATRIX_KIND_BAND && src->kind == BFM_MATRIX_KIND_BAND) {
		if (matrix->band.k != src->band.k) {
			return -1;
		}

		memcpy(matrix->band.data, src->band.data, src->m * (2 * src->band.k + 1) * sizeof *src->band.data);
		return 0;
	}

	if (matrix->kind == BFM_MATRIX_KIND_CSR && src->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_copy(matrix, src);
	}

	/* mixed kinds: element by element (matrix.c:426-438) */

	for (size_t i = 0; i < matrix->m; i++) {
		for (size_t j = 0; j < matrix->m; j++) {
			double const val = bfm_matrix_get(src, i, j);

			if (BFM_IS_NAN(val) || bfm_matrix_set(matrix, i, j, val) < 0) {
				return -1;
			}
		}
	}

	return 0;
}

size_t bfm_matrix_bandwidth(bfm_matrix_t* matrix) {
	if (matrix->kind == BFM_MATRIX_KIND_BAND) {
		return matrix->band.k;
	}

	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_bandwidth(matrix);
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL) {
		return (size_t) -1;
	}

	size_t k = 0;

	for (size_t i = 0; i < matrix->m; i++) { /* matrix.c:57-71: NaN counts as non-zero */
		for (size_t j = 0; j < matrix->m; j++) {
			size_t const d = i > j ? i - j : j - i;

			if (d > k && matrix->full.data[full_index(matrix, i, j)] != 0) {
				k = d;
			}
		}
	}

	return k;
}

/* ---- unpivoted LU (Doolittle, multipliers stored below the diagonal) --------------------------- */

/* FULL: matrix.c:73-121.  BAND: matrix.c:253-302 - same elimination restricted to the band. */
int bfm_matrix_lu(bfm_matrix_t* matrix) {
	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return 0; /* nothing to factor: bfm_matrix_lu_solve runs PCG on the GPU */
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL && matrix->kind != BFM_MATRIX_KIND_BAND) {
		return -1;
	}

	size_t const m = matrix->m;
	bool const band = matrix->kind == BFM_MATRIX_KIND_BAND;

	for (size_t p = 0; p + 1 < m; p++) {
		double const pivot = bfm_matrix_get(matrix, p, p);

		if (BFM_IS_NAN(pivot) || fabs(pivot) < BFM_PIVOT_EPS) {
			return -1;
		}

		size_t const reach = band ? BFM_MIN(p + matrix->band.k + 1, m) : m;

		for (size_t i = p + 1; i < reach; i++) {
			double factor = bfm_matrix_get(matrix, i, p);

			if (BFM_IS_NAN(factor)) {
				return -1;
			}

			factor /= pivot;

			if (bfm_matrix_set(matrix, i, p, factor) < 0) {
				return -1;
			}

			for (size_t j = p + 1; j < reach; j++) {
				double const above = bfm_matrix_get(matrix, p, j);

				if (BFM_IS_NAN(above) || bfm_matrix_add(matrix, i, j, -factor * above) < 0) {
					return -1;
				}
			}
		}
	}

	return 0;
}

/* forward (unit lower) then backward substitution, in place on vec.
 * FULL: matrix.c:123-174.  BAND: matrix.c:351-404. */
int bfm_matrix_lu_solve(bfm_matrix_t* matrix, bfm_vec_t* vec) {
	if (matrix->m != vec->n) {
		return -1;
	}

	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_solve(matrix, vec);
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL && matrix->kind != BFM_MATRIX_KIND_BAND) {
		return -1;
	}

	size_t const m = matrix->m;
	size_t const k = matrix->kind == BFM_MATRIX_KIND_BAND ? matrix->band.k : m;
	double* const y = vec->data;

	for (size_t p = 0; p < m; p++) {
		for (size_t i = p > k ? p - k : 0; i < p; i++) {
			double const val = bfm_matrix_get(matrix, p, i);

			if (BFM_IS_NAN(val)) {
				return -1;
			}

			y[p] -= val * y[i];
		}
	}

	for (size_t p = m; p-- > 0;) {
		size_t const reach = BFM_MIN(p + k + 1, m);

		for (size_t i = p + 1; i < reach; i++) {
			double const val = bfm_matrix_get(matrix, p, i);

			if (BFM_IS_NAN(val)) {
				return -1;
			}

			y[p] -= y[i] * val; /* operand order of matrix.c:390; the product is commutative */
		}

		double const pivot = bfm_matrix_get(matrix, p, p);

		if (BFM_IS_NAN(pivot) || (matrix->kind == BFM_MATRIX_KIND_BAND && pivot == 0)) {
			return -1;
		}

		y[p] /= pivot;
	}

	return 0;
}

int bfm_matrix_solve(bfm_matrix_t* matrix, bfm_vec_t* vec) {
	if (bfm_matrix_lu(matrix) < 0) {
		return -1;
	}

	return bfm_matrix_lu_solve(matrix, vec);
}
