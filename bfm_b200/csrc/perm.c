/*
 * perm.c - permutations and Reverse Cuthill-McKee (bfm/perm.h).
 *
 * The ordering must match the reference's bit for bit (BASELINE.json north_star), so the traversal
 * keeps every tie-break of reference perm.c:117-331 while dropping its O(n^2) dense scans:
 *   - degree of a DOF = number of numerically NON-ZERO entries in its row, diagonal included (:140-144)
 *   - each component starts from the unvisited DOF of smallest degree, the LAST one among equals (:222-233)
 *   - neighbours are gathered along the row in ascending index, marked visited when gathered (:257-272),
 *     then sorted by degree with a stable sort - glibc's qsort is a merge sort (:274) - and queued
 *   - DOFs are written into inv_perm back to front in pop order (:284); perm is its inverse (:297-299)
 * The traversal only needs "the non-zero columns of row i in ascending order", supplied by a
 * callback, so one driver serves dense (FULL) and sparse (CSR) matrices at O(nnz) cost.
 */
#include "internal.h"

#include <string.h>

int bfm_perm_create(bfm_perm_t* perm, bfm_state_t* state, size_t m) { /* perm.c:5-13 */
	memset(perm, 0, sizeof *perm);

	perm->state = state;
	perm->m = m;

	return 0;
}

int bfm_perm_destroy(bfm_perm_t* perm) { /* perm.c:15-24 */
	if (perm->has_perm) {
		perm->state->free(perm->perm);
		perm->state->free(perm->inv_perm);
	}

	return 0;
}

/* v'[p[i]] = v[i]  (perm.c:70-103) */
int bfm_perm_perm_vec(bfm_perm_t* perm, bfm_vec_t* vec, bool inv) {
	if (!perm->has_perm || vec->n != perm->m) {
		return -1;
	}

	size_t const* const p = inv ? perm->inv_perm : perm->perm;
	double* const old = perm->state->alloc(vec->n * sizeof *old);

	if (old == NULL) {
		return -1;
	}

	memcpy(old, vec->data, vec->n * sizeof *old);

	for (size_t i = 0; i < vec->n; i++) {
		vec->data[p[i]] = old[i];
	}

	perm->state->free(old);
	return 0;
}

/* A'[p[i]][p[j]] = A[i][j]  (perm.c:26-68).  FULL: physical; CSR: recorded as a logical renumbering. */
int bfm_perm_perm_matrix(bfm_perm_t* perm, bfm_matrix_t* matrix, bool inv) {
	if (!perm->has_perm || matrix->m != perm->m) {
		return -1;
	}

	size_t const* const p = inv ? perm->inv_perm : perm->perm;

	if (matrix->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_set_perm(matrix, p, perm->m);
	}

	if (matrix->kind != BFM_MATRIX_KIND_FULL) {
		return -1;
	}

	size_t const m = matrix->m;
	bfm_matrix_t old;

	if (bfm_matrix_full_create(&old, perm->state, matrix->major, m) < 0) {
		return -1;
	}

	bfm_matrix_copy(&old, matrix);

	for (size_t i = 0; i < m; i++) {
		for (size_t j = 0; j < m; j++) {
			bfm_matrix_set(matrix, p[i], p[j], bfm_matrix_get(&old, i, j));
		}
	}

	bfm_matrix_destroy(&old);
	return 0;
}

/* ---- RCM --------------------------------------------------------------------------------------- */

typedef struct {
	size_t dof;
	size_t deg;
} ranked_t;

/* the reference compares (int) deg differences (perm.c:110-115) */
static inline bool deg_after(ranked_t const* a, ranked_t const* b) {
	return (int) a->deg - (int) b->deg > 0;
}

/* stable: equal degrees keep their gather order */
static void sort_by_degree(ranked_t* v, ranked_t* tmp, size_t n) {
	if (n <= 24) {
		for (size_t i = 1; i < n; i++) {
			ranked_t const cur = v[i];
			size_t j = i;

			for (; j > 0 && deg_after(&v[j - 1], &cur); j--) {
				v[j] = v[j - 1];
			}

			v[j] = cur;
		}

		return;
	}

	size_t const half = n / 2;

	sort_by_degree(v, tmp, half);
	sort_by_degree(v + half, tmp, n - half);

	size_t l = 0, r = half, o = 0;

	while (l < half && r < n) {
		tmp[o++] = deg_after(&v[l], &v[r]) ? v[r++] : v[l++];
	}

	while (l < half) {
		tmp[o++] = v[l++];
	}

	while (r < n) {
		tmp[o++] = v[r++];
	}

	memcpy(v, tmp, n * sizeof *v);
}

int bfmi_rcm(bfm_perm_t* perm, size_t n, size_t max_row, bfmi_row_fn_t row_nnz, void* ctx) {
	bfm_state_t* const state = perm->state;
	int rv = -1;

	if (perm->m != n) {
		return -1;
	}

	size_t* const degs = state->alloc((n + 1) * sizeof *degs);
	size_t* const queue = state->alloc((n + 1) * sizeof *queue);
	bool* const visited = state->alloc((n + 1) * sizeof *visited);
	size_t* const row = state->alloc((max_row + 1) * sizeof *row);
	ranked_t* const found = state->alloc((max_row + 1) * sizeof *found);
	ranked_t* const tmp = state->alloc((max_row + 1) * sizeof *tmp);
	size_t* const bucket = state->alloc((max_row + 2) * sizeof *bucket);
	size_t* const by_degree = state->alloc((n + 1) * sizeof *by_degree);
	size_t* const inv_perm = state->alloc((n + 1) * sizeof *inv_perm);
	size_t* fwd = NULL;

	if (!degs || !queue || !visited || !row || !found || !tmp || !bucket || !by_degree || !inv_perm) {
		goto done;
	}

	memset(visited, 0, (n + 1) * sizeof *visited);
	memset(bucket, 0, (max_row + 2) * sizeof *bucket);

	for (size_t i = 0; i < n; i++) {
		degs[i] = row_nnz(i, row, ctx);
		bucket[degs[i] + 1]++;
	}

	/* start candidates ordered by (degree ascending, index DESCENDING): a counting sort filled from
	 * the highest index down.  A cursor over this list yields, at any time, exactly the DOF the
	 * reference's full rescan would pick, because visited DOFs only ever accumulate. */

	for (size_t d = 0; d <= max_row; d++) {
		bucket[d + 1] += bucket[d];
	}

	for (size_t i = n; i-- > 0;) {
		by_degree[bucket[degs[i]]++] = i;
	}

	size_t cursor = 0;
	size_t placed = 0;

	while (placed < n) {
		while (visited[by_degree[cursor]]) {
			cursor++;
		}

		size_t head = 0, tail = 0;

		queue[tail++] = by_degree[cursor];
		visited[by_degree[cursor]] = true; /* the reference marks it when popped; nothing can observe the gap */

		while (head != tail) {
			size_t const cur = queue[head++];
			size_t const cnt = row_nnz(cur, row, ctx);
			size_t n_found = 0;

			for (size_t t = 0; t < cnt; t++) {
				if (!visited[row[t]]) {
					visited[row[t]] = true;
					found[n_found++] = (ranked_t) {.dof = row[t], .deg = degs[row[t]]};
				}
			}

			sort_by_degree(found, tmp, n_found);

			for (size_t t = 0; t < n_found; t++) {
				queue[tail++] = found[t].dof;
			}

			inv_perm[n - ++placed] = cur;
		}
	}

	fwd = state->alloc((n + 1) * sizeof *fwd);

	if (fwd == NULL) {
		goto done;
	}

	for (size_t i = 0; i < n; i++) {
		fwd[inv_perm[i]] = i;
	}

	rv = 0;

done:

	state->free(degs);
	state->free(queue);
	state->free(visited);
	state->free(row);
	state->free(found);
	state->free(tmp);
	state->free(bucket);
	state->free(by_degree);

	if (rv < 0) {
		state->free(inv_perm);
		return -1;
	}

	perm->perm = fwd;
	perm->inv_perm = inv_perm;
	perm->has_perm = true;

	return 0;
}

static size_t full_row_nnz(size_t i, size_t* out, void* ctx) {
	bfm_matrix_t* const A = ctx;
	size_t cnt = 0;

	for (size_t j = 0; j < A->m; j++) {
		if (bfm_matrix_get(A, i, j) != 0) { /* `!!value`: NaN counts as an entry */
			out[cnt++] = j;
		}
	}

	return cnt;
}

int bfm_perm_rcm(bfm_perm_t* perm, bfm_matrix_t* A) {
	if (A->kind == BFM_MATRIX_KIND_CSR) {
		return bfmi_csr_rcm(perm, A);
	}

	if (A->kind != BFM_MATRIX_KIND_FULL && A->kind != BFM_MATRIX_KIND_BAND) {
		return -1;
	}

	return bfmi_rcm(perm, A->m, A->m, full_row_nnz, A);
}
