/*
 * bfm_oracle.c - sparse CPU restatement of libbfm's hot path (see bfm_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY.  Every routine cites the reference lines it restates; paths are
 * relative to the reference tree (libbfm/src/...).  Compile with -ffp-contract=off (oracle/Makefile):
 * the reference is plain IEEE-754 double arithmetic without fused multiply-add.
 */
#include "bfm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PIVOT_EPS 1e-20 /* bfm/math.h:7 */
#define ORC_MAX_POINTS 16

/* ------------------------------------------------------------------------------------------------
 * quadrature tables (rule.c:90-111).  The quad abscissa is the reference's truncated literal.
 * ---------------------------------------------------------------------------------------------- */

int orc_rule_gauss_legendre(int kind, double* weights, double* points) {
	double const sixth = 1. / 6;
	double const third = 1. / 3;
	double const s = 0.577350269189626;

	if (kind == 3) {
		double const p[3][2] = {{sixth, sixth}, {1 - third, sixth}, {sixth, 1 - third}};

		for (int g = 0; g < 3; g++) {
			weights[g] = sixth;
			points[2 * g + 0] = p[g][0];
			points[2 * g + 1] = p[g][1];
		}

		return 0;
	}

	if (kind == 4) {
		double const p[4][2] = {{-s, s}, {-s, -s}, {s, -s}, {s, s}};

		for (int g = 0; g < 4; g++) {
			weights[g] = 1;
			points[2 * g + 0] = p[g][0];
			points[2 * g + 1] = p[g][1];
		}

		return 0;
	}

	return -1;
}

/* ------------------------------------------------------------------------------------------------
 * shape functions (shape.c:3-38) and their derivatives (shape.c:40-85)
 * ---------------------------------------------------------------------------------------------- */

static void shape_phi(int kind, double const* pt, double* phi) {
	double const xsi = pt[0];
	double const eta = pt[1];

	if (kind == 3) {
		phi[0] = 1 - xsi - eta;
		phi[1] = xsi;
		phi[2] = eta;
	}

	else {
		phi[0] = (1 + xsi) * (1 + eta) / 4;
		phi[1] = (1 - xsi) * (1 + eta) / 4;
		phi[2] = (1 - xsi) * (1 - eta) / 4;
		phi[3] = (1 + xsi) * (1 - eta) / 4;
	}
}

static void shape_dphi(int kind, int wrt, double const* pt, double* d) {
	double const xsi = pt[0];
	double const eta = pt[1];

	if (kind == 3) {
		d[0] = -1;
		d[1] = wrt == 0 ? 1 : 0;
		d[2] = wrt == 0 ? 0 : 1;
	}

	else if (wrt == 0) {
		d[0] = (1 + eta) / 4;
		d[1] = (-1 - eta) / 4;
		d[2] = (-1 + eta) / 4;
		d[3] = (1 - eta) / 4;
	}

	else {
		d[0] = (1 + xsi) / 4;
		d[1] = (1 - xsi) / 4;
		d[2] = (-1 + xsi) / 4;
		d[3] = (-1 - xsi) / 4;
	}
}

/* ------------------------------------------------------------------------------------------------
 * sparse storage standing in for the dense matrix of system.c:5-34 / matrix.c:555-568
 * ---------------------------------------------------------------------------------------------- */

static int cmp_size(void const* a, void const* b) {
	size_t const x = *(size_t const*) a;
	size_t const y = *(size_t const*) b;
	return x < y ? -1 : x > y;
}

/* structural pattern: DOF (2a+r) couples with DOF (2c+s) whenever nodes a and c share an element */
static int build_pattern(orc_mesh_t const* mesh, orc_system_t* sys) {
	size_t const kind = (size_t) mesh->kind;
	size_t const nn = mesh->n_nodes;
	size_t const ne = mesh->n_elems;
	int rv = -1;

	size_t* n2e_ptr = calloc(nn + 2, sizeof *n2e_ptr);
	size_t* n2e = malloc((ne * kind + 1) * sizeof *n2e);
	size_t* nbr_ptr = calloc(nn + 1, sizeof *nbr_ptr);
	size_t* nbr = NULL;
	size_t* scratch = NULL;

	if (!n2e_ptr || !n2e || !nbr_ptr) {
		goto done;
	}

	for (size_t e = 0; e < ne; e++) {
		for (size_t j = 0; j < kind; j++) {
			size_t const a = mesh->elems[e * kind + j];

			if (a >= nn) {
				goto done;
			}

			n2e_ptr[a + 2]++;
		}
	}

	for (size_t a = 0; a < nn; a++) {
		n2e_ptr[a + 2] += n2e_ptr[a + 1];
	}

	for (size_t e = 0; e < ne; e++) {
		for (size_t j = 0; j < kind; j++) {
			n2e[n2e_ptr[mesh->elems[e * kind + j] + 1]++] = e;
		}
	}

	/* now n2e_ptr[a] .. n2e_ptr[a+1] lists the elements touching node a */

	size_t max_cand = 1;

	for (size_t a = 0; a < nn; a++) {
		size_t const c = (n2e_ptr[a + 1] - n2e_ptr[a]) * kind + 1;
		max_cand = c > max_cand ? c : max_cand;
	}

	scratch = malloc(max_cand * sizeof *scratch);

	if (!scratch) {
		goto done;
	}

	for (int pass = 0; pass < 2; pass++) {
		size_t total = 0;

		for (size_t a = 0; a < nn; a++) {
			size_t cnt = 0;

			scratch[cnt++] = a; /* the diagonal always exists in the dense matrix */

			for (size_t t = n2e_ptr[a]; t < n2e_ptr[a + 1]; t++) {
				for (size_t j = 0; j < kind; j++) {
					scratch[cnt++] = mesh->elems[n2e[t] * kind + j];
				}
			}

			qsort(scratch, cnt, sizeof *scratch, cmp_size);

			size_t uniq = 0;

			for (size_t t = 0; t < cnt; t++) {
				if (t == 0 || scratch[t] != scratch[t - 1]) {
					if (pass == 1) {
						nbr[total + uniq] = scratch[t];
					}

					uniq++;
				}
			}

			if (pass == 0) {
				nbr_ptr[a + 1] = uniq;
			}

			total += uniq;
		}

		if (pass == 0) {
			for (size_t a = 0; a < nn; a++) {
				nbr_ptr[a + 1] += nbr_ptr[a];
			}

			nbr = malloc((total + 1) * sizeof *nbr);

			if (!nbr) {
				goto done;
			}
		}
	}

	size_t const n = 2 * nn;
	size_t const nnz = 4 * nbr_ptr[nn];

	sys->n = n;
	sys->rowptr = malloc((n + 1) * sizeof *sys->rowptr);
	sys->col = malloc((nnz + 1) * sizeof *sys->col);
	sys->val = calloc(nnz + 1, sizeof *sys->val);
	sys->b = calloc(n + 1, sizeof *sys->b);

	if (!sys->rowptr || !sys->col || !sys->val || !sys->b) {
		goto done;
	}

	size_t pos = 0;

	for (size_t a = 0; a < nn; a++) {
		for (size_t r = 0; r < 2; r++) {
			sys->rowptr[2 * a + r] = pos;

			for (size_t t = nbr_ptr[a]; t < nbr_ptr[a + 1]; t++) {
				sys->col[pos++] = 2 * nbr[t] + 0;
				sys->col[pos++] = 2 * nbr[t] + 1;
			}
		}
	}

	sys->rowptr[n] = pos;
	rv = 0;

done:

	free(n2e_ptr);
	free(n2e);
	free(nbr_ptr);
	free(nbr);
	free(scratch);

	return rv;
}

/* position of (i, j) in the row structure, or (size_t) -1 */
static size_t find_entry(orc_system_t const* sys, size_t i, size_t j) {
	size_t lo = sys->rowptr[i];
	size_t hi = sys->rowptr[i + 1];

	while (lo < hi) {
		size_t const mid = lo + (hi - lo) / 2;

		if (sys->col[mid] < j) {
			lo = mid + 1;
		}

		else {
			hi = mid;
		}
	}

	if (lo < sys->rowptr[i + 1] && sys->col[lo] == j) {
		return lo;
	}

	return (size_t) -1;
}

/* bfm_matrix_add on the dense matrix (matrix.c:46-55) */
static void sys_add(orc_system_t* sys, size_t i, size_t j, double v) {
	size_t const at = find_entry(sys, i, j);

	if (at != (size_t) -1) {
		sys->val[at] += v;
	}
}

double orc_system_get(orc_system_t const* sys, size_t i, size_t j) {
	if (i >= sys->n || j >= sys->n) {
		return 0. / 0.; /* matrix.c:26-28 */
	}

	size_t const at = find_entry(sys, i, j);
	return at == (size_t) -1 ? 0 : sys->val[at];
}

void orc_system_destroy(orc_system_t* sys) {
	free(sys->rowptr);
	free(sys->col);
	free(sys->val);
	free(sys->b);
	memset(sys, 0, sizeof *sys);
}

/* ------------------------------------------------------------------------------------------------
 * element loop: get_elem (system.c:93-107), fill_elasticity_elem (system.c:109-227),
 * fill_axisymmetric_elem (system.c:229-356)
 * ---------------------------------------------------------------------------------------------- */

static void fill_elem(orc_problem_t const* p, orc_system_t* sys, size_t e, int axisym, double a, double b, double c) {
	orc_mesh_t const* const mesh = &p->mesh;
	size_t const kind = (size_t) mesh->kind;

	size_t map[4];
	double x[4];
	double y[4];

	for (size_t j = 0; j < kind; j++) { /* system.c:99-106 */
		map[j] = mesh->elems[kind * e + j];
		x[j] = mesh->coords[map[j] * 2 + 0];
		y[j] = mesh->coords[map[j] * 2 + 1];
	}

	for (size_t g = 0; g < p->n_points; g++) { /* system.c:136 */
		double const weight = p->weights[g];
		double const* const pt = &p->points[2 * g];

		double phi[4];
		double dphi_dxsi[4];
		double dphi_deta[4];

		shape_phi(mesh->kind, pt, phi);
		shape_dphi(mesh->kind, 0, pt, dphi_dxsi);
		shape_dphi(mesh->kind, 1, pt, dphi_deta);

		/* jacobian (system.c:164-176, :290-305) */

		double dx_dxsi = 0;
		double dx_deta = 0;
		double dy_dxsi = 0;
		double dy_deta = 0;
		double r = 0;

		for (size_t j = 0; j < kind; j++) {
			dx_dxsi += x[j] * dphi_dxsi[j];
			dx_deta += x[j] * dphi_deta[j];
			dy_dxsi += y[j] * dphi_dxsi[j];
			dy_deta += y[j] * dphi_deta[j];
			r += x[j] * phi[j];
		}

		double const det_J = fabs(dx_dxsi * dy_deta - dx_deta * dy_dxsi);

		/* physical gradients (system.c:183-186) */

		double dphi_dx[4];
		double dphi_dy[4];

		for (size_t j = 0; j < kind; j++) {
			dphi_dx[j] = (dphi_dxsi[j] * dy_deta - dphi_deta[j] * dy_dxsi) / det_J;
			dphi_dy[j] = (dphi_deta[j] * dx_dxsi - dphi_dxsi[j] * dx_deta) / det_J;
		}

		/* load vector (system.c:190-203, :319-332) */

		for (size_t j = 0; j < kind; j++) {
			size_t const row = 2 * map[j];

			for (size_t k = 0; k < p->n_forces; k++) {
				double const* const F = &p->forces[(k * mesh->n_nodes + map[j]) * 2];

				if (!axisym) {
					sys->b[row + 0] += det_J * weight * F[0] * p->rho * phi[j];
					sys->b[row + 1] += det_J * weight * F[1] * p->rho * phi[j];
				}

				else {
					sys->b[row + 0] += det_J * weight * F[0] * p->rho * phi[j] * r;
					sys->b[row + 1] += det_J * weight * F[1] * p->rho * phi[j] * r;
				}
			}
		}

		/* stiffness (system.c:207-223, :336-352) */

		for (size_t j = 0; j < kind; j++) {
			size_t const row = 2 * map[j];

			for (size_t k = 0; k < kind; k++) {
				size_t const col = 2 * map[k];

				double f_11, f_12, f_21, f_22;

				if (!axisym) {
					f_11 = a * dphi_dx[j] * dphi_dx[k] + c * dphi_dy[j] * dphi_dy[k];
					f_12 = b * dphi_dx[j] * dphi_dy[k] + c * dphi_dy[j] * dphi_dx[k];
					f_21 = b * dphi_dy[j] * dphi_dx[k] + c * dphi_dx[j] * dphi_dy[k];
					f_22 = a * dphi_dy[j] * dphi_dy[k] + c * dphi_dx[j] * dphi_dx[k];
				}

				else {
					f_11 = a * dphi_dx[j] * dphi_dx[k] * r + c * dphi_dy[j] * dphi_dy[k] * r + phi[j] * (b * dphi_dx[k] + a * phi[k] / r) + dphi_dx[j] * b * phi[k];
					f_12 = b * dphi_dx[j] * dphi_dy[k] * r + c * dphi_dy[j] * dphi_dx[k] * r + phi[j] * b * dphi_dy[k];
					f_21 = b * dphi_dy[j] * dphi_dx[k] * r + c * dphi_dx[j] * dphi_dy[k] * r + dphi_dy[j] * b * phi[k];
					f_22 = a * dphi_dy[j] * dphi_dy[k] * r + c * dphi_dx[j] * dphi_dx[k];
				}

				sys_add(sys, row + 0, col + 0, det_J * weight * f_11);
				sys_add(sys, row + 0, col + 1, det_J * weight * f_12);
				sys_add(sys, row + 1, col + 0, det_J * weight * f_21);
				sys_add(sys, row + 1, col + 1, det_J * weight * f_22);
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------
 * boundary conditions
 * ---------------------------------------------------------------------------------------------- */

/* apply_constraint (system.c:358-374).  The dense sweeps visit every row/column; outside the
 * structural pattern they compute b[i] -= value * 0 and store zeros over zeros, so only the
 * pattern needs to be walked.  The pattern is structurally symmetric, hence the rows holding
 * column `dof` are exactly the columns of row `dof`. */
static void apply_constraint(orc_system_t* sys, size_t dof, double value) {
	for (size_t t = sys->rowptr[dof]; t < sys->rowptr[dof + 1]; t++) {
		size_t const i = sys->col[t];
		size_t const at = find_entry(sys, i, dof);

		if (at == (size_t) -1) {
			continue;
		}

		sys->b[i] -= value * sys->val[at];
		sys->val[at] = 0;
	}

	for (size_t t = sys->rowptr[dof]; t < sys->rowptr[dof + 1]; t++) {
		sys->val[t] = 0;
	}

	sys->val[find_entry(sys, dof, dof)] = 1;
	sys->b[dof] = value;
}

/* apply_dirichlet (system.c:376-386) */
static void apply_dirichlet(orc_problem_t const* p, orc_system_t* sys, orc_condition_t const* cond) {
	size_t const shift = cond->kind == 0 ? 0 : 1;

	for (size_t j = 0; j < p->mesh.n_nodes; j++) {
		if (cond->nodes[j]) {
			apply_constraint(sys, j * 2 + shift, cond->value);
		}
	}
}

/* apply_dirichlet_normal_tangent (system.c:388-425).  pow(x, 2) is x * x (what gcc -O2 emits). */
static void apply_dirichlet_nt(orc_problem_t const* p, orc_system_t* sys, orc_condition_t const* cond) {
	orc_mesh_t const* const mesh = &p->mesh;
	int const tangent = cond->kind == 7;

	for (size_t i = 0; i < mesh->n_nodes; i++) {
		if (!cond->nodes[i]) {
			continue;
		}

		double tx = 0;
		double ty = 0;

		for (size_t j = 0; j < mesh->n_edges; j++) {
			if (mesh->edge_elems[2 * j + 1] != -1) {
				continue;
			}

			size_t n2;

			if (mesh->edge_nodes[2 * j + 0] == i) {
				n2 = mesh->edge_nodes[2 * j + 1];
			}

			else if (mesh->edge_nodes[2 * j + 1] == i) {
				n2 = mesh->edge_nodes[2 * j + 0];
			}

			else {
				continue;
			}

			double const dx = mesh->coords[i * 2 + 0] - mesh->coords[n2 * 2 + 0];
			double const dy = mesh->coords[i * 2 + 1] - mesh->coords[n2 * 2 + 1];
			double const length = sqrt(dx * dx + dy * dy);

			tx += dx / length / 2;
			ty += dy / length / 2;
		}

		apply_constraint(sys, 2 * i + 0, cond->value * (tangent ? tx : -ty));
		apply_constraint(sys, 2 * i + 1, cond->value * (tangent ? ty : tx));
	}
}

/* Neumann X / Y edge loads (system.c:477-498; axisymmetric weighting :586-617) */
static void apply_neumann_xy(orc_problem_t const* p, orc_system_t* sys, orc_condition_t const* cond, int axisym) {
	orc_mesh_t const* const mesh = &p->mesh;
	size_t const shift = cond->kind == 2 ? 0 : 1;

	for (size_t j = 0; j < mesh->n_edges; j++) {
		size_t const n1 = mesh->edge_nodes[2 * j + 0];
		size_t const n2 = mesh->edge_nodes[2 * j + 1];

		if (!cond->nodes[n1] || !cond->nodes[n2]) {
			continue;
		}

		double const dx = mesh->coords[n1 * 2 + 0] - mesh->coords[n2 * 2 + 0];
		double const dy = mesh->coords[n1 * 2 + 1] - mesh->coords[n2 * 2 + 1];
		double const c2 = dx * dx + dy * dy;
		double const jacobian = sqrt(c2) / 2;

		if (!axisym) {
			sys->b[n1 * 2 + shift] += jacobian * cond->value;
			sys->b[n2 * 2 + shift] += jacobian * cond->value;
		}

		else {
			double const r1 =
				mesh->coords[n1 * 2 + 0] * (1 - 1 / sqrt(3)) / 2 +
				mesh->coords[n1 * 2 + 1] * (1 + 1 / sqrt(3)) / 2;

			double const r2 =
				mesh->coords[n2 * 2 + 0] * (1 - 1 / sqrt(3)) / 2 +
				mesh->coords[n2 * 2 + 1] * (1 + 1 / sqrt(3)) / 2;

			double const fac = r1 + r2;

			sys->b[n1 * 2 + shift] += fac * jacobian * cond->value;
			sys->b[n2 * 2 + shift] += fac * jacobian * cond->value;
		}
	}
}

/* Neumann normal / tangent edge loads (system.c:500-522) */
static void apply_neumann_nt(orc_problem_t const* p, orc_system_t* sys, orc_condition_t const* cond) {
	orc_mesh_t const* const mesh = &p->mesh;
	int const tangent = cond->kind == 5;

	for (size_t j = 0; j < mesh->n_edges; j++) {
		size_t const n1 = mesh->edge_nodes[2 * j + 0];
		size_t const n2 = mesh->edge_nodes[2 * j + 1];

		if (!cond->nodes[n1] || !cond->nodes[n2]) {
			continue;
		}

		double const dx = mesh->coords[n1 * 2 + 0] - mesh->coords[n2 * 2 + 0];
		double const dy = mesh->coords[n1 * 2 + 1] - mesh->coords[n2 * 2 + 1];

		sys->b[n1 * 2 + 0] += 0.5 * cond->value * (tangent ? dx : -dy);
		sys->b[n1 * 2 + 1] += 0.5 * cond->value * (tangent ? dy : dx);
		sys->b[n2 * 2 + 0] += 0.5 * cond->value * (tangent ? dx : -dy);
		sys->b[n2 * 2 + 1] += 0.5 * cond->value * (tangent ? dy : dx);
	}
}

/* ------------------------------------------------------------------------------------------------
 * system creation: create_planar (system.c:427-529), axisymmetric (system.c:539-622)
 * ---------------------------------------------------------------------------------------------- */

static int create(orc_problem_t const* p, orc_system_t* sys, int with_bcs) {
	memset(sys, 0, sizeof *sys);

	if (p->mesh.kind != 3 && p->mesh.kind != 4) { /* system.c:440-442 */
		return -1;
	}

	if (p->sim_kind < 1 || p->sim_kind > 3 || p->n_points > ORC_MAX_POINTS) {
		return -1;
	}

	if (build_pattern(&p->mesh, sys) < 0) {
		orc_system_destroy(sys);
		return -1;
	}

	int const axisym = p->sim_kind == 3;
	int const stress = p->sim_kind == 2;

	double const E = p->E;
	double const nu = p->nu;

	double a, b;

	if (axisym) { /* system.c:242-244, note the '*' where the planar path divides */
		a = E * (1 - nu) / (1 + nu) * (1 - 2 * nu);
		b = E * nu / (1 + nu) / (1 - 2 * nu);
	}

	else { /* system.c:454-456 */
		a = !stress ? E * (1 - nu) / (1 + nu) / (1 - 2 * nu) : E / (1 - nu * nu);
		b = !stress ? E * nu / (1 + nu) / (1 - 2 * nu) : E * nu / (1 - nu * nu);
	}

	double const c = E / (2 * (1 + nu)); /* system.c:458 */

	for (size_t e = 0; e < p->mesh.n_elems; e++) { /* system.c:460-466 */
		fill_elem(p, sys, e, axisym, a, b, c);
	}

	if (!with_bcs) {
		return 0;
	}

	for (size_t i = 0; i < p->n_conditions; i++) { /* system.c:470-526, :575-619 */
		orc_condition_t const* const cond = &p->conditions[i];

		if (cond->kind == 0 || cond->kind == 1) {
			apply_dirichlet(p, sys, cond);
		}

		else if (cond->kind == 2 || cond->kind == 3) {
			apply_neumann_xy(p, sys, cond, axisym);
		}

		else if ((cond->kind == 4 || cond->kind == 5) && !axisym) {
			apply_neumann_nt(p, sys, cond);
		}

		else if (cond->kind == 6 || cond->kind == 7) {
			apply_dirichlet_nt(p, sys, cond);
		}
	}

	return 0;
}

int orc_system_create(orc_problem_t const* p, orc_system_t* sys) {
	return create(p, sys, 1);
}

int orc_system_assemble(orc_problem_t const* p, orc_system_t* sys) {
	return create(p, sys, 0);
}

/* ------------------------------------------------------------------------------------------------
 * Reverse Cuthill-McKee on the numeric pattern (perm.c:117-331)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
	size_t i;
	size_t deg;
} rcm_node_t;

/* glibc's qsort (perm.c:274) is a stable merge sort for these sizes; the comparison is on degree
 * only (perm.c:110-115), so equal degrees keep their gather order (ascending index). */
static void stable_sort_by_deg(rcm_node_t* v, size_t cnt, rcm_node_t* tmp) {
	if (cnt < 2) {
		return;
	}

	if (cnt <= 16) {
		for (size_t i = 1; i < cnt; i++) {
			rcm_node_t const cur = v[i];
			size_t j = i;

			while (j > 0 && (int) v[j - 1].deg - (int) cur.deg > 0) {
				v[j] = v[j - 1];
				j--;
			}

			v[j] = cur;
		}

		return;
	}

	size_t const half = cnt / 2;

	stable_sort_by_deg(v, half, tmp);
	stable_sort_by_deg(v + half, cnt - half, tmp);

	size_t l = 0, r = half, o = 0;

	while (l < half && r < cnt) {
		tmp[o++] = ((int) v[l].deg - (int) v[r].deg <= 0) ? v[l++] : v[r++];
	}

	while (l < half) {
		tmp[o++] = v[l++];
	}

	while (r < cnt) {
		tmp[o++] = v[r++];
	}

	memcpy(v, tmp, cnt * sizeof *v);
}

/* start-node order: smallest degree first, LAST index among equals (perm.c:222-233) */
static int cmp_start(void const* _a, void const* _b) {
	rcm_node_t const* const a = _a;
	rcm_node_t const* const b = _b;

	if (a->deg != b->deg) {
		return a->deg < b->deg ? -1 : 1;
	}

	return a->i > b->i ? -1 : a->i < b->i;
}

int orc_rcm(orc_system_t const* sys, size_t* perm, size_t* inv_perm) {
	size_t const n = sys->n;
	int rv = -1;

	size_t* degs = calloc(n + 1, sizeof *degs);
	size_t* queue = malloc((n + 1) * sizeof *queue);
	unsigned char* visited = calloc(n + 1, 1);
	rcm_node_t* to_sort = malloc((n + 1) * sizeof *to_sort);
	rcm_node_t* tmp = malloc((n + 1) * sizeof *tmp);
	rcm_node_t* starts = malloc((n + 1) * sizeof *starts);

	if (!degs || !queue || !visited || !to_sort || !tmp || !starts) {
		goto done;
	}

	for (size_t i = 0; i < n; i++) { /* perm.c:140-144: the diagonal counts */
		for (size_t t = sys->rowptr[i]; t < sys->rowptr[i + 1]; t++) {
			degs[i] += sys->val[t] != 0;
		}

		starts[i].i = i;
		starts[i].deg = degs[i];
	}

	qsort(starts, n, sizeof *starts, cmp_start);

	size_t cursor = 0;
	size_t unvisited = n;
	size_t head = 0, tail = 0;

	for (size_t p = 0; unvisited && p < n;) { /* perm.c:215 */
		while (visited[starts[cursor].i]) {
			cursor++;
		}

		queue[tail++] = starts[cursor].i; /* perm.c:237 - NOT yet marked visited */

		while (head != tail) {
			size_t const cur = queue[head++];

			if (!visited[cur]) { /* perm.c:245-248 */
				visited[cur] = 1;
				unvisited--;
			}

			size_t cnt = 0;

			for (size_t t = sys->rowptr[cur]; t < sys->rowptr[cur + 1]; t++) { /* perm.c:257-272 */
				size_t const i = sys->col[t];

				if (visited[i] || sys->val[t] == 0) {
					continue;
				}

				to_sort[cnt].i = i;
				to_sort[cnt].deg = degs[i];
				cnt++;

				visited[i] = 1;
			}

			/* NB the reference marks gathered neighbours visited without decrementing its
			 * unvisited counter (perm.c:271 vs :245-248); the counter only reaches zero through
			 * the pops, which is what the loop above reproduces. */

			stable_sort_by_deg(to_sort, cnt, tmp);

			for (size_t t = 0; t < cnt; t++) {
				queue[tail++] = to_sort[t].i;
			}

			inv_perm[n - p++ - 1] = cur; /* perm.c:284 */
		}
	}

	for (size_t i = 0; i < n; i++) { /* perm.c:297-299 */
		perm[inv_perm[i]] = i;
	}

	rv = 0;

done:

	free(degs);
	free(queue);
	free(visited);
	free(to_sort);
	free(tmp);
	free(starts);

	return rv;
}

/* ------------------------------------------------------------------------------------------------
 * renumber -> band -> LU -> substitution (system.c:44-81, matrix.c:57-71, :253-302, :351-404)
 * ---------------------------------------------------------------------------------------------- */

size_t orc_bandwidth(orc_system_t const* sys, size_t const* perm) {
	size_t k = 0;

	for (size_t i = 0; i < sys->n; i++) {
		for (size_t t = sys->rowptr[i]; t < sys->rowptr[i + 1]; t++) {
			if (sys->val[t] == 0) { /* matrix.c:62: `!value` */
				continue;
			}

			size_t const pi = perm ? perm[i] : i;
			size_t const pj = perm ? perm[sys->col[t]] : sys->col[t];
			size_t const d = pi > pj ? pi - pj : pj - pi;

			k = d > k ? d : k;
		}
	}

	return k;
}

int orc_band_solve(orc_system_t const* sys, size_t const* perm, double* x) {
	size_t const m = sys->n;
	size_t const k = orc_bandwidth(sys, perm);
	int rv = -1;

	/* band buffer, element (i, j) at j + i * 2k (matrix.c:208, :570-584) */

	double* band = calloc(m * (2 * k + 1) + 1, sizeof *band);
	double* y = malloc((m + 1) * sizeof *y);

	if (!band || !y) {
		goto done;
	}

#define BAND(i, j) band[(j) + (i) * 2 * k]

	/* A'[perm[i]][perm[j]] = A[i][j] (perm.c:57-65), b'[perm[i]] = b[i] (perm.c:97-100) */

	for (size_t i = 0; i < m; i++) {
		size_t const pi = perm ? perm[i] : i;

		for (size_t t = sys->rowptr[i]; t < sys->rowptr[i + 1]; t++) {
			size_t const pj = perm ? perm[sys->col[t]] : sys->col[t];
			size_t const d = pi > pj ? pi - pj : pj - pi;

			if (d <= k) { /* beyond the band the value is an exact zero by construction */
				BAND(pi, pj) = sys->val[t];
			}
		}

		y[pi] = sys->b[i];
	}

	/* matrix_band_lu (matrix.c:253-302) */

	for (size_t p = 0; m && p < m - 1; p++) {
		double const pivot = BAND(p, p);

		if (pivot != pivot || fabs(pivot) < ORC_PIVOT_EPS) {
			goto done;
		}

		size_t const len = p + k + 1 < m ? p + k + 1 : m;

		for (size_t i = p + 1; i < len; i++) {
			double below = BAND(i, p);

			if (below != below) {
				goto done;
			}

			below /= pivot;
			BAND(i, p) = below;

			for (size_t j = p + 1; j < len; j++) {
				double const v = BAND(p, j);

				if (v != v) {
					goto done;
				}

				BAND(i, j) += -below * v;
			}
		}
	}

	/* matrix_band_lu_solve (matrix.c:351-404) */

	for (ssize_t p = 0; p < (ssize_t) m; p++) {
		ssize_t const lo = p - (ssize_t) k > 0 ? p - (ssize_t) k : 0;

		for (ssize_t i = lo; i < p; i++) {
			double const v = BAND(p, i);

			if (v != v) {
				goto done;
			}

			y[p] -= v * y[i];
		}
	}

	for (ssize_t p = (ssize_t) m - 1; p >= 0; p--) {
		ssize_t const hi = p + (ssize_t) k + 1 < (ssize_t) m ? p + (ssize_t) k + 1 : (ssize_t) m;

		for (ssize_t i = p + 1; i < hi; i++) {
			double const v = BAND(p, i);

			if (v != v) {
				goto done;
			}

			y[p] -= y[i] * v;
		}

		double const pivot = BAND(p, p);

		if (pivot != pivot || !pivot) {
			goto done;
		}

		y[p] /= pivot;
	}

#undef BAND

	/* bfm_perm_perm_vec(inv = true): x[inv_perm[i]] = y[i], i.e. x[i] = y[perm[i]] (sim.c:123) */

	for (size_t i = 0; i < m; i++) {
		x[i] = y[perm ? perm[i] : i];
	}

	rv = 0;

done:

	free(band);
	free(y);

	return rv;
}

void orc_spmv(orc_system_t const* sys, double const* x, double* y) {
	for (size_t i = 0; i < sys->n; i++) {
		double acc = 0;

		for (size_t t = sys->rowptr[i]; t < sys->rowptr[i + 1]; t++) {
			acc += sys->val[t] * x[sys->col[t]];
		}

		y[i] = acc;
	}
}

/* run_elasticity for one instance (sim.c:103-135) */
int orc_run(orc_problem_t const* p, double* effects) {
	orc_system_t sys;
	int rv = -1;

	if (orc_system_create(p, &sys) < 0) {
		return -1;
	}

	size_t* perm = malloc((sys.n + 1) * sizeof *perm);
	size_t* inv_perm = malloc((sys.n + 1) * sizeof *inv_perm);

	if (!perm || !inv_perm) {
		goto done;
	}

	if (orc_rcm(&sys, perm, inv_perm) < 0) {
		goto done;
	}

	/* the reference ignores bfm_matrix_solve's status (sim.c:122); we report it */

	rv = orc_band_solve(&sys, perm, effects);

done:

	free(perm);
	free(inv_perm);
	orc_system_destroy(&sys);

	return rv;
}
