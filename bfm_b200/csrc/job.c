/*
 * job.c - the hot path's driver: what the reference's run_elasticity (sim.c:103-135) and
 * create_planar / bfm_system_create_axisymmetric_strain (system.c:427-529, :539-622) do for one
 * instance, re-staged for the GPU:
 *
 *   create    host: plan (cached per mesh), shape tables, material constants, BC work lists
 *   upload    H2D:  coordinates, per-node force tables (FUNKY forces only), BC lists
 *   assemble  GPU:  k_assemble, then one k_bc_dirichlet / k_bc_add per boundary condition, in
 *                   instance condition order (system.c:470-526)
 *   solve     GPU:  FP64 PCG
 *   download  D2H:  displacements -> instance->effects (sim.c:127-131)
 *
 * bfm_sim_run chains the five; bfm_system_create_planar_* stop after `assemble` and hand the matrix
 * back as a CSR-kind bfm_matrix_t.
 */
#include "internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- BC work lists ----------------------------------------------------------------------------- */

typedef enum {
	OP_DIRICHLET, /* apply_constraint on an ascending DOF list */
	OP_ADD,       /* ordered additions to the right-hand side */
} op_kind_t;

typedef struct {
	op_kind_t kind;

	/* OP_DIRICHLET: constrained DOFs + values, block rows to revisit */
	int32_t n_dofs;
	int32_t* dofs;
	double* vals;
	int32_t n_rows;
	int32_t* rows;

	/* OP_ADD: groups of additions per DOF, in application order inside a group */
	int32_t n_groups;
	int32_t* group_dof;
	int32_t* group_ptr;
	double* add;

	/* device mirrors */
	int32_t *d_dofs, *d_rows, *d_group_dof, *d_group_ptr;
	double *d_vals, *d_add;
} bc_op_t;

typedef struct {
	bfm_instance_t* instance;
	bfm_mesh_t const* mesh;
	size_t node_off; /* first node of this system in the combined numbering (multiple of 32) */
} batch_sys_t;

struct bfmx_job {
	bfm_state_t* state;
	bfm_sim_kind_t kind;
	bfm_instance_t* instance;
	bfm_mesh_t const* gmesh; /* the instance's mesh (global numbering) */
	bfm_mesh_t const* mesh;  /* what this rank assembles: gmesh, or the local mesh of its partition */

	bfmi_renum_t* renum;     /* bfm_sim_run on a mesh numbered without locality: the internally renumbered copy the job works on (renumber.c); else NULL */
	double* d_xp;            /* ... and the solution back in the caller's numbering */

	bfmi_part_t* part;       /* NULL on one GPU */
	bfmi_plan_t* plan;       /* of `mesh` */
	bfmg_pattern_t pat;      /* plan->dev with the owned row range */
	bfmg_halo_t halo;
	double* d_xg;            /* multi-GPU: the gathered global solution */
	bfmi_coarse_t* coarse;   /* the solver's single coarse level (NULL: small mesh, disabled, or superseded by hier) */
	bfmi_hier_t* hier;       /* the solver's aggregation hierarchy (multilevel preconditioner), or NULL */
	bfmg_asm_tables_t tab;

	double* h_nforce; /* [n_forces][nb][2], FUNKY forces sampled at the nodes */

	size_t n_ops;
	bc_op_t* ops;

	double* d_coords;
	double* d_nforce;
	double* d_val;
	double* d_b;
	double* d_x;
	int32_t* d_stamp;
	double* d_cval;

	size_t table_forces;          /* forces per node in d_nforce (FUNKY tables) */

	/* batch of independent systems sharing one pattern (bfmx_job_create_batch); n_sys == 0 otherwise */
	int32_t n_sys;
	batch_sys_t* sys;
	bfm_mesh_t combined;          /* the systems' meshes as disconnected components, each padded to 32 nodes */
	bfmg_asm_tables_t* h_tabs;    /* [n_sys] */
	bfmg_asm_tables_t* d_tabs;
	int32_t* h_slice_tab;         /* [n_slices] system of each slice */
	int32_t* d_slice_tab;
	bfmg_batch_range_t* ranges;   /* [n_sys] */
	bfmg_batch_status_t* status;  /* [n_sys], filled by solve */

	bool uploaded, assembled, solved;
	bfmx_stats_t stats;
};

static bfmx_stats_t last_stats;

void bfmx_publish_stats(bfmx_stats_t const* stats) {
	last_stats = *stats;
}

int bfmx_last_stats(bfmx_stats_t* out) {
	*out = last_stats;
	return 0;
}

int bfmx_device_available(void) {
	return bfmg_available();
}

char const* bfmx_device_error(void) {
	return bfmg_last_error();
}

int bfmx_device_sync(void) {
	return bfmg_sync();
}

int bfmx_timer_start(int slot) {
	return bfmg_timer_start(slot);
}

float bfmx_timer_stop(int slot) {
	return bfmg_timer_stop(slot);
}

size_t bfmx_kernel_launches(void) {
	return bfmg_launch_count();
}

int bfmx_device_sm_count(void) {
	return bfmg_sm_count();
}

int bfmx_dist_unique_id(void* id) {
	return bfmg_dist_unique_id(id);
}

int bfmx_dist_init(int rank, int world, void const* id) {
	return bfmg_dist_init(rank, world, id);
}

int bfmx_dist_finalize(void) {
	return bfmg_dist_finalize();
}

int bfmx_dist_rank(void) {
	return bfmg_dist_rank();
}

int bfmx_dist_world(void) {
	return bfmg_dist_world();
}

char const* bfmx_dist_peer_memory_status(void) {
	return bfmg_dist_p2p_status();
}

static int batch_download(bfmx_job_t* job);

int bfmx_batch_max_nodes(void) {
	return bfmg_batch_max_rows();
}

static double now_ms(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ---- shape tables, material constants, forces --------------------------------------------------- */

/* T: the tables of one system.  FUNKY forces are sampled at the nodes of `mesh` into
 * nforce[(k * table_nodes + node_off + a) * 2 + d]; *nforce is allocated on first need with room for
 * table_forces forces over table_nodes nodes (one system: its own counts; a batch: the batch's) */
static int fill_tables(bfm_state_t* state, bfm_sim_kind_t sim_kind, bfm_instance_t* instance, bfm_mesh_t const* mesh, size_t n_forces, bfm_force_t** forces, bfmg_asm_tables_t* T, double** nforce, size_t table_forces, size_t table_nodes, size_t node_off) {
	bfm_obj_t* const obj = instance->obj;
	bfm_material_t const* const material = obj->material;
	bfm_rule_t* const rule = obj->rule;
	bfm_shape_t* const shape = &rule->shape;
	size_t const kind = mesh->kind;

	memset(T, 0, sizeof *T);

	if (rule->n_points > BFMG_MAX_POINTS) {
		return BFMI_FAIL(state, "integration rules with more than %d points are not supported on the GPU path", BFMG_MAX_POINTS);
	}

	if (n_forces > BFMG_MAX_FORCES) {
		return BFMI_FAIL(state, "more than %d body forces are not supported on the GPU path", BFMG_MAX_FORCES);
	}

	T->kind = (int32_t) kind;
	T->n_points = (int32_t) rule->n_points;
	T->axisym = sim_kind == BFM_SIM_KIND_AXISYMMETRIC_STRAIN;
	T->rho = material->rho;

	/* the rule's own shape functions, evaluated on the host exactly as the reference does per element
	 * and integration point (system.c:151-158); the kernel only reads the table */

	T->grad_const = 1;

	for (size_t g = 0; g < rule->n_points; g++) {
		double phi[6] = {0}, dxsi[6] = {0}, deta[6] = {0};

		shape->phi(shape, rule->points[g], phi);
		shape->dphi(shape, 0, rule->points[g], dxsi);
		shape->dphi(shape, 1, rule->points[g], deta);

		T->weight[g] = rule->weights[g];

		for (size_t j = 0; j < kind; j++) {
			T->phi[g][j] = phi[j];
			T->dxsi[g][j] = dxsi[j];
			T->deta[g][j] = deta[j];

			if (memcmp(&T->dxsi[g][j], &T->dxsi[0][j], sizeof(double)) != 0 || memcmp(&T->deta[g][j], &T->deta[0][j], sizeof(double)) != 0) {
				T->grad_const = 0;
			}
		}
	}

	/* elasticity constants, association as written in the reference */

	double const E = material->E;
	double const nu = material->nu;

	if (sim_kind == BFM_SIM_KIND_AXISYMMETRIC_STRAIN) { /* system.c:242-244 (note '* (1 - 2 nu)') */
		T->a = E * (1 - nu) / (1 + nu) * (1 - 2 * nu);
		T->b = E * nu / (1 + nu) / (1 - 2 * nu);
	}

	else if (sim_kind == BFM_SIM_KIND_PLANAR_STRAIN) { /* system.c:454-456 */
		T->a = E * (1 - nu) / (1 + nu) / (1 - 2 * nu);
		T->b = E * nu / (1 + nu) / (1 - 2 * nu);
	}

	else {
		T->a = E / (1 - nu * nu);
		T->b = E * nu / (1 - nu * nu);
	}

	T->c = E / (2 * (1 + nu));

	/* forces: constants when none is FUNKY, otherwise every force becomes a per-node table */

	T->n_forces = (int32_t) n_forces;

	for (size_t k = 0; k < n_forces; k++) {
		if (forces[k]->dim != 2) {
			return BFMI_FAIL(state, "force %zu has dimension %zu, the mesh has 2 (the reference would use stale values here)", k, forces[k]->dim);
		}

		if (forces[k]->kind == BFM_FORCE_KIND_FUNKY) {
			T->forces_per_node = 1;
		}
	}

	if (!T->forces_per_node) {
		for (size_t k = 0; k < n_forces; k++) {
			if (forces[k]->kind == BFM_FORCE_KIND_LINEAR) {
				T->const_force[k][0] = forces[k]->linear.force.data[0];
				T->const_force[k][1] = forces[k]->linear.force.data[1];
			}
		}

		return 0;
	}

	size_t const nb = mesh->n_nodes;

	if (*nforce == NULL) {
		*nforce = calloc(table_forces * table_nodes * 2 + 1, sizeof **nforce);
	}

	if (*nforce == NULL) {
		return -1;
	}

	bfm_vec_t pos, out;

	if (bfm_vec_create(&pos, state, 2) < 0 || bfm_vec_create(&out, state, 2) < 0) {
		return -1;
	}

	for (size_t k = 0; k < n_forces; k++) {
		for (size_t a = 0; a < nb; a++) {
			pos.data[0] = mesh->coords[2 * a + 0];
			pos.data[1] = mesh->coords[2 * a + 1];

			bfm_force_eval(forces[k], &pos, &out); /* status ignored, as in system.c:198 */

			(*nforce)[(k * table_nodes + node_off + a) * 2 + 0] = out.data[0];
			(*nforce)[(k * table_nodes + node_off + a) * 2 + 1] = out.data[1];
		}
	}

	bfm_vec_destroy(&pos);
	bfm_vec_destroy(&out);

	return 0;
}

/* ---- boundary conditions -> ordered work lists ---------------------------------------------------- */

static int cmp_i32(void const* a, void const* b) {
	int32_t const x = *(int32_t const*) a;
	int32_t const y = *(int32_t const*) b;
	return x < y ? -1 : x > y;
}

/* block rows holding a column of one of the constrained nodes = their graph neighbours */
static int affected_rows(bfmi_plan_t const* plan, bc_op_t* op) {
	size_t cap = 0;

	for (int32_t i = 0; i < op->n_dofs; i++) {
		if (i == 0 || op->dofs[i] / 2 != op->dofs[i - 1] / 2) {
			cap += (size_t) plan->row_len[op->dofs[i] / 2];
		}
	}

	op->rows = malloc((cap + 1) * sizeof *op->rows);

	if (op->rows == NULL) {
		return -1;
	}

	size_t cnt = 0;

	for (int32_t i = 0; i < op->n_dofs; i++) {
		int32_t const a = op->dofs[i] / 2;

		if (i != 0 && a == op->dofs[i - 1] / 2) {
			continue;
		}

		for (int32_t t = 0; t < plan->row_len[a]; t++) {
			op->rows[cnt++] = plan->scol[(int64_t) plan->slice_off[a / 32] + (int64_t) t * 32 + a % 32];
		}
	}

	qsort(op->rows, cnt, sizeof *op->rows, cmp_i32);

	size_t uniq = 0;

	for (size_t i = 0; i < cnt; i++) {
		if (i == 0 || op->rows[i] != op->rows[i - 1]) {
			op->rows[uniq++] = op->rows[i];
		}
	}

	op->n_rows = (int32_t) uniq;
	return 0;
}

/* apply_dirichlet (system.c:376-386) */
static int op_dirichlet_xy(bfm_mesh_t const* gmesh, bfm_condition_t const* cond, bc_op_t* op) {
	size_t const nn = gmesh->n_nodes;
	int32_t const shift = cond->kind == BFM_CONDITION_KIND_DIRICHLET_X ? 0 : 1;
	size_t count = 0, cap = 1024;

	/* the mask is one byte per node and almost empty on a large mesh (25 MB for a few thousand clamped nodes at
	 * 50 M DOF): eight bytes at a time, one pass */

	op->kind = OP_DIRICHLET;
	op->dofs = malloc(cap * sizeof *op->dofs);

	if (op->dofs == NULL) {
		return -1;
	}

	for (size_t j = 0; j < nn;) {
		if (j + 8 <= nn) {
			uint64_t word;

			memcpy(&word, &cond->nodes[j], sizeof word);

			if (word == 0) {
				j += 8;
				continue;
			}
		}

		size_t const end = j + 8 <= nn ? j + 8 : nn;

		for (; j < end; j++) {
			if (cond->nodes[j]) {
				if (count == cap) {
					int32_t* const grown = realloc(op->dofs, cap * 2 * sizeof *op->dofs);

					if (grown == NULL) {
						return -1;
					}

					op->dofs = grown;
					cap *= 2;
				}

				op->dofs[count++] = (int32_t) (2 * j) + shift;
			}
		}
	}

	op->vals = malloc((count + 1) * sizeof *op->vals);

	if (op->vals == NULL) {
		return -1;
	}

	for (size_t i = 0; i < count; i++) {
		op->vals[i] = cond->value;
	}

	op->n_dofs = (int32_t) count;

	return 0;
}

/* apply_dirichlet_normal_tangent (system.c:388-425): tangent = sum over the node's boundary edges, in
 * edge order, of (x_i - x_other) / length / 2; pow(d, 2) is d * d, as gcc -O2 compiles it */
static int op_dirichlet_nt(bfm_mesh_t const* mesh, bfm_condition_t const* cond, bc_op_t* op) {
	size_t const nn = mesh->n_nodes;
	bool const tangent = cond->kind == BFM_CONDITION_KIND_DIRICHLET_TANGENT;

	double* const tx = calloc(nn + 1, sizeof *tx);
	double* const ty = calloc(nn + 1, sizeof *ty);
	size_t count = 0;

	for (size_t j = 0; j < nn; j++) {
		count += cond->nodes[j];
	}

	op->kind = OP_DIRICHLET;
	op->dofs = malloc((2 * count + 1) * sizeof *op->dofs);
	op->vals = malloc((2 * count + 1) * sizeof *op->vals);

	if (tx == NULL || ty == NULL || op->dofs == NULL || op->vals == NULL) {
		free(tx);
		free(ty);
		return -1;
	}

	/* one pass over the edges in ascending order feeds every node's sum in the reference's order */

	for (size_t j = 0; j < mesh->n_edges; j++) {
		bfm_edge_t const* const edge = &mesh->edges[j];

		if (edge->elems[1] != -1) {
			continue;
		}

		for (int side = 0; side < 2; side++) {
			size_t const i = edge->nodes[side];
			size_t const other = edge->nodes[1 - side];

			if (side == 1 && edge->nodes[0] == edge->nodes[1]) {
				break; /* degenerate edge: the reference's else-if takes the first branch only */
			}

			if (i >= nn || other >= nn || !cond->nodes[i]) {
				continue;
			}

			double const dx = mesh->coords[i * 2 + 0] - mesh->coords[other * 2 + 0];
			double const dy = mesh->coords[i * 2 + 1] - mesh->coords[other * 2 + 1];
			double const length = sqrt(dx * dx + dy * dy);

			tx[i] += dx / length / 2;
			ty[i] += dy / length / 2;
		}
	}

	for (size_t i = 0; i < nn; i++) {
		if (!cond->nodes[i]) {
			continue;
		}

		op->dofs[op->n_dofs] = (int32_t) (2 * i);
		op->vals[op->n_dofs++] = cond->value * (tangent ? tx[i] : -ty[i]);

		op->dofs[op->n_dofs] = (int32_t) (2 * i + 1);
		op->vals[op->n_dofs++] = cond->value * (tangent ? ty[i] : tx[i]);
	}

	free(tx);
	free(ty);

	return 0;
}

typedef struct {
	int32_t dof;
	int32_t seq;
	double val;
} pending_add_t;

static int cmp_pending(void const* a, void const* b) {
	pending_add_t const* const x = a;
	pending_add_t const* const y = b;

	if (x->dof != y->dof) {
		return x->dof < y->dof ? -1 : 1;
	}

	return x->seq < y->seq ? -1 : x->seq > y->seq;
}

/* Neumann loads (system.c:477-522; axisymmetric weighting :586-617): every mesh edge with both end
 * nodes in the mask adds to the right-hand side, in edge order */
static int op_neumann(bfm_mesh_t const* mesh, bool axisym, bfm_condition_t const* cond, bc_op_t* op) {
	bool const xy = cond->kind == BFM_CONDITION_KIND_NEUMANN_X || cond->kind == BFM_CONDITION_KIND_NEUMANN_Y;

	op->kind = OP_ADD;

	if (!xy && axisym) {
		return 0; /* the axisymmetric BC loop has no normal/tangent Neumann branch */
	}

	pending_add_t* const pend = malloc((4 * mesh->n_edges + 1) * sizeof *pend);

	if (pend == NULL) {
		return -1;
	}

	int32_t n = 0;

	for (size_t j = 0; j < mesh->n_edges; j++) {
		size_t const n1 = mesh->edges[j].nodes[0];
		size_t const n2 = mesh->edges[j].nodes[1];

		if (n1 >= mesh->n_nodes || n2 >= mesh->n_nodes || !cond->nodes[n1] || !cond->nodes[n2]) {
			continue;
		}

		double const dx = mesh->coords[n1 * 2 + 0] - mesh->coords[n2 * 2 + 0];
		double const dy = mesh->coords[n1 * 2 + 1] - mesh->coords[n2 * 2 + 1];

		if (xy) {
			int32_t const shift = cond->kind == BFM_CONDITION_KIND_NEUMANN_X ? 0 : 1;
			double const jacobian = sqrt(dx * dx + dy * dy) / 2;
			double load = jacobian * cond->value;

			if (axisym) {
				double const r1 = mesh->coords[n1 * 2 + 0] * (1 - 1 / sqrt(3)) / 2 + mesh->coords[n1 * 2 + 1] * (1 + 1 / sqrt(3)) / 2;
				double const r2 = mesh->coords[n2 * 2 + 0] * (1 - 1 / sqrt(3)) / 2 + mesh->coords[n2 * 2 + 1] * (1 + 1 / sqrt(3)) / 2;
				double const fac = r1 + r2;

				load = fac * jacobian * cond->value;
			}

			pend[n] = (pending_add_t) {(int32_t) (n1 * 2) + shift, n, load}, n++;
			pend[n] = (pending_add_t) {(int32_t) (n2 * 2) + shift, n, load}, n++;
		}

		else {
			bool const tangent = cond->kind == BFM_CONDITION_KIND_NEUMANN_TANGENT;
			double const lx = 0.5 * cond->value * (tangent ? dx : -dy);
			double const ly = 0.5 * cond->value * (tangent ? dy : dx);

			pend[n] = (pending_add_t) {(int32_t) (n1 * 2 + 0), n, lx}, n++;
			pend[n] = (pending_add_t) {(int32_t) (n1 * 2 + 1), n, ly}, n++;
			pend[n] = (pending_add_t) {(int32_t) (n2 * 2 + 0), n, lx}, n++;
			pend[n] = (pending_add_t) {(int32_t) (n2 * 2 + 1), n, ly}, n++;
		}
	}

	qsort(pend, (size_t) n, sizeof *pend, cmp_pending);

	op->group_dof = malloc(((size_t) n + 1) * sizeof *op->group_dof);
	op->group_ptr = malloc(((size_t) n + 2) * sizeof *op->group_ptr);
	op->add = malloc(((size_t) n + 1) * sizeof *op->add);

	if (op->group_dof == NULL || op->group_ptr == NULL || op->add == NULL) {
		free(pend);
		return -1;
	}

	for (int32_t i = 0; i < n; i++) {
		if (i == 0 || pend[i].dof != pend[i - 1].dof) {
			op->group_dof[op->n_groups] = pend[i].dof;
			op->group_ptr[op->n_groups++] = i;
		}

		op->add[i] = pend[i].val;
	}

	op->group_ptr[op->n_groups] = n;

	free(pend);
	return 0;
}

/* Multi-GPU: the lists above are in global DOF numbers; keep the DOFs of local nodes (owned and ghost -
 * a ghost column's Dirichlet value feeds the owned rows' right-hand side) and renumber them.  The local
 * numbering is monotone in the global one, so ascending order - the reference's application order - holds. */
static void localize_op(bfmi_part_t const* part, bc_op_t* op) {
	if (op->kind == OP_DIRICHLET) {
		int32_t kept = 0;

		for (int32_t i = 0; i < op->n_dofs; i++) {
			int32_t const l = bfmi_part_local(part, (size_t) op->dofs[i] / 2);

			if (l >= 0) {
				op->dofs[kept] = 2 * l + op->dofs[i] % 2;
				op->vals[kept++] = op->vals[i];
			}
		}

		op->n_dofs = kept;
		return;
	}

	/* right-hand-side additions: only owned rows matter */

	int32_t kept_groups = 0;
	int32_t kept_adds = 0;

	for (int32_t g = 0; g < op->n_groups; g++) {
		int32_t const l = bfmi_part_local(part, (size_t) op->group_dof[g] / 2);
		int32_t const beg = op->group_ptr[g];
		int32_t const end = op->group_ptr[g + 1];

		if (l < part->own_begin || l >= part->own_end) {
			continue;
		}

		op->group_dof[kept_groups] = 2 * l + op->group_dof[g] % 2;
		op->group_ptr[kept_groups++] = kept_adds;

		for (int32_t t = beg; t < end; t++) {
			op->add[kept_adds++] = op->add[t];
		}
	}

	if (op->n_groups > 0) {
		op->group_ptr[kept_groups] = kept_adds;
	}

	op->n_groups = kept_groups;
}

/* the lists above are in the caller's DOF numbers; the job works on the internal numbering of renumber.c.  Dirichlet
 * lists are put back in ascending order of the new numbers (k_bc_dirichlet and affected_rows expect ascending DOFs;
 * which DOFs are constrained and to what value does not change); the additions of a group keep their order */
typedef struct {
	int32_t dof;
	int32_t at;
} dof_at_t;

static int cmp_dof_at(void const* a, void const* b) {
	dof_at_t const* const x = a;
	dof_at_t const* const y = b;
	return x->dof < y->dof ? -1 : x->dof > y->dof;
}

static int renumber_op(bfmi_renum_t const* renum, bc_op_t* op) {
	if (op->kind != OP_DIRICHLET) {
		for (int32_t g = 0; g < op->n_groups; g++) {
			op->group_dof[g] = 2 * renum->to_new[op->group_dof[g] / 2] + op->group_dof[g] % 2;
		}

		return 0;
	}

	dof_at_t* const order = malloc(((size_t) op->n_dofs + 1) * sizeof *order);
	double* const vals = malloc(((size_t) op->n_dofs + 1) * sizeof *vals);

	if (order == NULL || vals == NULL) {
		free(order), free(vals);
		return -1;
	}

	for (int32_t i = 0; i < op->n_dofs; i++) {
		order[i].dof = 2 * renum->to_new[op->dofs[i] / 2] + op->dofs[i] % 2;
		order[i].at = i;
	}

	qsort(order, (size_t) op->n_dofs, sizeof *order, cmp_dof_at);

	for (int32_t i = 0; i < op->n_dofs; i++) {
		op->dofs[i] = order[i].dof;
		vals[i] = op->vals[order[i].at];
	}

	free(order);
	free(op->vals);
	op->vals = vals;

	return 0;
}

/* one instance's conditions, in instance order, as work lists in the DOF numbering of `gmesh` */
static int build_ops_for(bfm_mesh_t const* gmesh, bfm_sim_kind_t sim_kind, bfm_instance_t const* instance, bc_op_t** out_ops, size_t* out_n) {
	bool const axisym = sim_kind == BFM_SIM_KIND_AXISYMMETRIC_STRAIN;
	bc_op_t* const ops = calloc(instance->n_conditions + 1, sizeof *ops);
	size_t n_ops = 0;

	*out_ops = ops;
	*out_n = 0;

	if (ops == NULL) {
		return -1;
	}

	for (size_t i = 0; i < instance->n_conditions; i++) {
		bfm_condition_t const* const cond = instance->conditions[i];
		bc_op_t* const op = &ops[n_ops];
		int rv = 0;

		switch (cond->kind) {
		case BFM_CONDITION_KIND_DIRICHLET_X:
		case BFM_CONDITION_KIND_DIRICHLET_Y:
			rv = op_dirichlet_xy(gmesh, cond, op);
			break;

		case BFM_CONDITION_KIND_DIRICHLET_NORMAL:
		case BFM_CONDITION_KIND_DIRICHLET_TANGENT:
			rv = op_dirichlet_nt(gmesh, cond, op);
			break;

		case BFM_CONDITION_KIND_NEUMANN_X:
		case BFM_CONDITION_KIND_NEUMANN_Y:
			rv = op_neumann(gmesh, axisym, cond, op);
			break;

		case BFM_CONDITION_KIND_NEUMANN_NORMAL:
		case BFM_CONDITION_KIND_NEUMANN_TANGENT:
			if (axisym) {
				continue;
			}

			rv = op_neumann(gmesh, axisym, cond, op);
			break;

		default:
			continue; /* unknown kinds fall through every branch of the reference's BC loop */
		}

		*out_n = ++n_ops; /* counted before the check so that a failed list is still released */

		if (rv < 0) {
			return -1;
		}
	}

	return 0;
}

static void free_ops(bc_op_t* ops, size_t n_ops) {
	for (size_t i = 0; i < n_ops; i++) {
		bc_op_t* const op = &ops[i];

		free(op->dofs), free(op->vals), free(op->rows);
		free(op->group_dof), free(op->group_ptr), free(op->add);

		bfmg_free(op->d_dofs), bfmg_free(op->d_vals), bfmg_free(op->d_rows);
		bfmg_free(op->d_group_dof), bfmg_free(op->d_group_ptr), bfmg_free(op->d_add);
	}

	free(ops);
}

static int build_ops(bfmx_job_t* job) {
	if (build_ops_for(job->gmesh, job->kind, job->instance, &job->ops, &job->n_ops) < 0) {
		return -1;
	}

	for (size_t i = 0; i < job->n_ops; i++) {
		bc_op_t* const op = &job->ops[i];

		if (job->renum != NULL && renumber_op(job->renum, op) < 0) {
			return -1;
		}

		if (job->part != NULL) {
			localize_op(job->part, op);
		}

		if (op->kind == OP_DIRICHLET && affected_rows(job->plan, op) < 0) {
			return -1;
		}
	}

	return 0;
}

/* ---- job life cycle -------------------------------------------------------------------------------- */

/* a small mesh on one GPU: the whole PCG runs inside one CTA (batch.cu) instead of three launches per
 * iteration; BFM_ONE_CTA=0 forces the general path */
static bool takes_one_cta(bfmx_job_t const* job) {
	char const* const env = getenv("BFM_ONE_CTA");

	return job->n_sys == 0 && job->part == NULL && job->plan->nb <= bfmg_batch_max_rows() && (env == NULL || atoi(env) != 0);
}

/* device buffers of a job whose plan, tables and work lists are ready */
static int job_alloc_device(bfmx_job_t* job, size_t table_forces) {
	bfm_state_t* const state = job->state;
	size_t const nb = (size_t) job->plan->nb;
	size_t const n_forces = table_forces;

	if (
		bfmg_alloc((void**) &job->d_coords, nb * 2 * sizeof(double)) < 0 ||
		bfmg_alloc((void**) &job->d_val, (size_t) job->plan->n_slots * 4 * sizeof(double)) < 0 ||
		bfmg_alloc((void**) &job->d_b, nb * 2 * sizeof(double)) < 0 ||
		bfmg_alloc((void**) &job->d_x, nb * 2 * sizeof(double)) < 0 ||
		(job->h_nforce != NULL && bfmg_alloc((void**) &job->d_nforce, n_forces * nb * 2 * sizeof(double)) < 0)
	) {
		BFMI_FAIL(state, "device allocation failed: %s", bfmg_last_error());
		return -1;
	}

	bool any_dirichlet = false;

	for (size_t i = 0; i < job->n_ops; i++) {
		bc_op_t* const op = &job->ops[i];
		int rv = 0;

		if (op->kind == OP_DIRICHLET) {
			any_dirichlet = true;

			rv |= bfmg_alloc((void**) &op->d_dofs, ((size_t) op->n_dofs + 1) * sizeof(int32_t));
			rv |= bfmg_alloc((void**) &op->d_vals, ((size_t) op->n_dofs + 1) * sizeof(double));
			rv |= bfmg_alloc((void**) &op->d_rows, ((size_t) op->n_rows + 1) * sizeof(int32_t));
		}

		else if (op->n_groups > 0) {
			rv |= bfmg_alloc((void**) &op->d_group_dof, ((size_t) op->n_groups + 1) * sizeof(int32_t));
			rv |= bfmg_alloc((void**) &op->d_group_ptr, ((size_t) op->n_groups + 2) * sizeof(int32_t));
			rv |= bfmg_alloc((void**) &op->d_add, ((size_t) op->group_ptr[op->n_groups] + 1) * sizeof(double));
		}

		if (rv < 0) {
			BFMI_FAIL(state, "device allocation failed: %s", bfmg_last_error());
			return -1;
		}
	}

	if (any_dirichlet && (bfmg_alloc((void**) &job->d_stamp, nb * 2 * sizeof(int32_t)) < 0 || bfmg_alloc((void**) &job->d_cval, nb * 2 * sizeof(double)) < 0)) {
		BFMI_FAIL(state, "device allocation failed: %s", bfmg_last_error());
		return -1;
	}

	return 0;
}

static int job_create(bfmx_job_t** out, bfm_state_t* state, bfm_sim_kind_t kind, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces, bool partitioned) {
	bfm_mesh_t* const mesh = instance->obj->mesh;

	*out = NULL;

	if (mesh->dim != 2 || (mesh->kind != BFM_ELEM_KIND_SIMPLEX && mesh->kind != BFM_ELEM_KIND_QUAD)) {
		return -1; /* system.c:436-442, :547-553 */
	}

	if (kind != BFM_SIM_KIND_PLANAR_STRAIN && kind != BFM_SIM_KIND_PLANAR_STRESS && kind != BFM_SIM_KIND_AXISYMMETRIC_STRAIN) {
		return -1;
	}

	if (!bfmg_available()) {
		return BFMI_FAIL(state, "%s", bfmg_last_error());
	}

	bfmx_job_t* const job = calloc(1, sizeof *job);

	if (job == NULL) {
		return -1;
	}

	job->state = state;
	job->kind = kind;
	job->instance = instance;
	job->gmesh = mesh;
	job->mesh = mesh;

	double const t0 = now_ms();
	bool const verbose = getenv("BFM_JOB_VERBOSE") != NULL;
	double t_mark = t0;

#define MARK(what) do { if (verbose) { double const now_ = now_ms(); fprintf(stderr, "[job] %-28s %8.2f ms\n", (what), now_ - t_mark); t_mark = now_; } } while (0)

	/* bfm_sim_run only (the public matrix API keeps the caller's numbering): the hash of the connectivity keys every
	 * per-mesh cache below.  Several GPUs: every rank holds the GLOBAL connectivity (1.2 GB at 50 M DOF), so each
	 * hashes its world-th of it and the parts are combined over the communicator */

	uint64_t hash = 0;
	bfm_mesh_t const* wmesh = mesh; /* the mesh the job works on, before any partition */

	if (partitioned) {
		if (bfmg_dist_world() == 1) {
			hash = bfmi_mesh_hash(mesh);
		}

		else {
			int const world = bfmg_dist_world();
			uint64_t const mine = bfmi_mesh_hash_part(mesh, bfmg_dist_rank(), world);
			double parts[2] = {(double) (uint32_t) mine, (double) (uint32_t) (mine >> 32)}; /* exact in a double */
			double all[2 * BFMG_DIST_MAX_RANKS];
			double *d_send = NULL, *d_recv = NULL;

			if (
				bfmg_alloc((void**) &d_send, sizeof parts) < 0 || bfmg_alloc((void**) &d_recv, sizeof all) < 0 ||
				bfmg_upload(d_send, parts, sizeof parts) < 0 || bfmg_dist_allgather_f64(d_send, d_recv, 2) < 0 ||
				bfmg_download(all, d_recv, (size_t) world * sizeof parts) < 0
			) {
				bfmg_free(d_send);
				bfmg_free(d_recv);
				BFMI_FAIL(state, "exchanging the mesh hash failed: %s", bfmg_last_error());
				goto fail;
			}

			bfmg_free(d_send);
			bfmg_free(d_recv);

			for (int r = 0; r < world; r++) {
				hash ^= (uint64_t) all[2 * r] | (uint64_t) all[2 * r + 1] << 32;
			}
		}

		MARK("connectivity hash");

		/* a numbering without locality (renumber.c): from here on the job works on an internally renumbered copy of
		 * the mesh - same elements, same order - and hands the displacements back in the caller's numbering */

		job->renum = bfmi_renum_for_mesh(mesh, hash);

		if (job->renum != NULL) {
			wmesh = &job->renum->mesh;
			hash = job->renum->mesh_hash;
			job->mesh = wmesh;

			if (bfmi_renum_upload(job->renum) < 0 || bfmg_alloc((void**) &job->d_xp, mesh->n_nodes * 2 * sizeof(double)) < 0) {
				BFMI_FAIL(state, "device allocation failed: %s", bfmg_last_error());
				goto fail;
			}
		}

		MARK("numbering (cache look-up)");
	}

	/* several GPUs: this rank assembles and solves the local mesh of its row block (partition.c) */

	if (partitioned && bfmg_dist_world() > 1) {
		job->part = bfmi_part_for_mesh(state, wmesh, hash, bfmg_dist_rank(), bfmg_dist_world());

		if (job->part == NULL) {
			goto fail;
		}

		job->mesh = &job->part->local;
	}

	MARK("partition (cache look-up)");

	job->plan = partitioned && job->part == NULL ? bfmi_plan_for_mesh_hashed(state, job->mesh, hash) : bfmi_plan_for_mesh(state, job->mesh);

	if (job->plan == NULL) {
		goto fail;
	}

	MARK("plan (cache look-up)");

	bool const fresh = !job->plan->accounted; /* built (and mirrored) by this very call: its time and bytes go into the stats once */

	job->plan->accounted = true;

	if (bfmi_plan_upload(state, job->plan) < 0) {
		goto fail;
	}

	if (fresh) {
		job->stats.ms_plan = (float) (now_ms() - t0);

		job->stats.h2d_bytes += job->plan->h2d_bytes;
		job->stats.d2h_bytes += job->plan->d2h_bytes;
	}

	MARK("plan upload");

	if (fill_tables(state, kind, instance, job->mesh, n_forces, forces, &job->tab, &job->h_nforce, n_forces, job->mesh->n_nodes, 0) < 0) {
		goto fail;
	}

	MARK("shape / force tables");

	if (build_ops(job) < 0) {
		goto fail;
	}

	MARK("boundary-condition work lists");

	job->pat = job->plan->dev;

	if (job->part != NULL) {
		bfmi_part_t* const part = job->part;

		job->pat.row_lo = part->own_begin;
		job->pat.row_hi = part->own_end;

		if (part->d_send_idx == NULL && part->n_send > 0) {
			if (bfmg_alloc((void**) &part->d_send_idx, (size_t) part->n_send * sizeof(int32_t)) < 0 || bfmg_upload(part->d_send_idx, part->send_idx, (size_t) part->n_send * sizeof(int32_t)) < 0) {
				BFMI_FAIL(state, "uploading the halo plan failed: %s", bfmg_last_error());
				goto fail;
			}

			job->stats.h2d_bytes += (size_t) part->n_send * sizeof(int32_t);
		}

		job->halo.n_nbr = part->n_nbr;
		job->halo.nbr = part->nbr;
		job->halo.recv_begin = part->recv_begin;
		job->halo.recv_count = part->recv_count;
		job->halo.send_ptr = part->send_ptr;
		job->halo.n_send = part->n_send;
		job->halo.d_send_idx = part->d_send_idx;

		if (bfmg_alloc((void**) &job->d_xg, mesh->n_nodes * 2 * sizeof(double)) < 0) {
			BFMI_FAIL(state, "device allocation failed: %s", bfmg_last_error());
			goto fail;
		}
	}

	job->table_forces = n_forces;

	if (job_alloc_device(job, n_forces) < 0) {
		goto fail;
	}

	MARK("device buffers");

	/* coarse level of the solver for meshes the one-CTA path does not take.  How many aggregates: iterations fall
	 * like 1 / sqrt(n_agg) while inverting E grows like n_agg^3 and applying E^-1 like n_agg^2 per iteration;
	 * minimising  c sqrt(n / n_agg) (a n + b n_agg^2) + d n_agg^3  with the measured constants gives
	 * n_agg ~ 3.4 n^(3/7): ~700 at 0.25 M nodes, ~2300 at 4 M, ~5000 at 25 M (measured there: 4.65 s with 2093
	 * aggregates, 3.83 s with 3850; the optimum is flat).  At least 32 nodes per aggregate; at most 5120 (E^-1 is
	 * 1.9 GB then), 2048 when several GPUs exchange over NCCL and therefore each invert all of E.
	 * BFM_COARSE_AGGREGATES overrides; 0 switches the coarse level off. */

	/* the multilevel preconditioner (hier.c, mg.cuh) supersedes that single level; BFM_MG=0 keeps the single level.
	 * On several GPUs it needs the exchanges over NVLink peer memory, and building the hierarchy is a collective
	 * call (every rank gets here: bfm_sim_run is collective on a partitioned job). */

	if (!takes_one_cta(job) && (job->part == NULL || bfmg_dist_p2p_status()[0] == 0)) {
		char const* const env = getenv("BFM_MG");

		if (env == NULL || atoi(env) != 0) {
			double const t_h = now_ms();

			job->hier = bfmi_hier_for_plan(job->plan, job->mesh->coords, job->part);

			if (job->hier != NULL) {
				bool const fresh_hier = !job->hier->on_device;

				if (bfmi_hier_upload(job->hier, &job->plan->dev, &job->stats.h2d_bytes) < 0) {
					BFMI_FAIL(state, "uploading the aggregation hierarchy failed: %s", bfmg_last_error());
					goto fail;
				}

				if (fresh_hier) {
					job->stats.ms_plan += (float) (now_ms() - t_h);
				}
			}
		}
	}

	if (!takes_one_cta(job) && job->hier == NULL) {
		char const* const env = getenv("BFM_COARSE_AGGREGATES");
		int64_t target = (int64_t) floor(3.4 * pow((double) mesh->n_nodes, 3.0 / 7.0) + 0.5);
		int64_t const cap = job->part != NULL && bfmg_dist_p2p_status()[0] != 0 ? 2048 : 5120;

		target = target > (int64_t) (mesh->n_nodes / 32) ? (int64_t) (mesh->n_nodes / 32) : target;
		target = target > cap ? cap : target;

		if (env != NULL) {
			target = atoll(env);
		}

		if (target >= 4) {
			bool none;

			job->coarse = bfmi_coarse_for_mesh(state, wmesh, job->plan->elems_hash, job->part, (int32_t) target, &none);

			if (job->coarse != NULL && bfmi_coarse_upload(job->coarse) < 0) {
				BFMI_FAIL(state, "uploading the coarse level failed: %s", bfmg_last_error());
				goto fail;
			}
		}
	}

	MARK("hierarchy (cache look-up)");
#undef MARK

	job->stats.n_dofs = 2 * mesh->n_nodes;
	job->stats.n_dofs_owned = 2 * (size_t) (job->pat.row_hi - job->pat.row_lo);
	job->stats.n_ranks = job->part != NULL ? (size_t) job->part->world : 1;
	job->stats.halo_bytes_per_exchange = job->part != NULL ? (size_t) job->part->n_send * 16 : 0;
	job->stats.n_blocks = (size_t) job->plan->n_blocks;
	job->stats.n_slots = (size_t) job->plan->n_slots;

	*out = job;
	return 0;

fail:

	bfmx_job_destroy(job);
	return -1;
}

int bfmx_job_create(bfmx_job_t** job, bfm_sim_t* sim, size_t instance_index) {
	if (instance_index >= sim->n_instances) {
		return -1;
	}

	return job_create(job, sim->state, sim->kind, sim->instances[instance_index], sim->n_forces, sim->forces, true);
}

/* ---- batches of small systems (BASELINE.json configs[4]; kernels in batch.cu) ---------------------------
 *
 * All systems of the batch become ONE mesh of disconnected components - system s occupies the node range
 * [node_off[s], node_off[s] + n_nodes[s]), node_off a multiple of 32 so that no SELL-32 slice straddles
 * two systems (the padding nodes are isolated: an empty diagonal block, x = 0).  Node and element order
 * inside a system are kept, so each system is assembled with the reference's accumulation order, bit for
 * bit as if it were alone.  Material, rule and forces differ per system: tables[s].  Conditions are
 * applied in instance order per system; the c-th conditions of all systems are merged into one work list
 * (systems are independent, so only the order inside a system matters). */

static void shift_op(bc_op_t* op, size_t node_off) {
	for (int32_t i = 0; i < op->n_dofs; i++) {
		op->dofs[i] += (int32_t) (2 * node_off);
	}

	for (int32_t g = 0; g < op->n_groups; g++) {
		op->group_dof[g] += (int32_t) (2 * node_off);
	}
}

/* dst (zeroed, same kind) += src: lists are appended; systems come in ascending node order, so DOFs stay ascending */
static int append_op(bc_op_t* dst, bc_op_t const* src) {
	if (src->kind == OP_DIRICHLET) {
		if (src->n_dofs == 0) {
			return 0;
		}

		int32_t* const dofs = realloc(dst->dofs, ((size_t) dst->n_dofs + src->n_dofs + 1) * sizeof *dofs);
		double* const vals = realloc(dst->vals, ((size_t) dst->n_dofs + src->n_dofs + 1) * sizeof *vals);

		dst->dofs = dofs != NULL ? dofs : dst->dofs;
		dst->vals = vals != NULL ? vals : dst->vals;

		if (dofs == NULL || vals == NULL) {
			return -1;
		}

		memcpy(dst->dofs + dst->n_dofs, src->dofs, (size_t) src->n_dofs * sizeof *dofs);
		memcpy(dst->vals + dst->n_dofs, src->vals, (size_t) src->n_dofs * sizeof *vals);
		dst->n_dofs += src->n_dofs;

		return 0;
	}

	if (src->n_groups == 0) {
		return 0;
	}

	int32_t const old_adds = dst->n_groups > 0 ? dst->group_ptr[dst->n_groups] : 0;
	int32_t const new_adds = src->group_ptr[src->n_groups];

	int32_t* const gd = realloc(dst->group_dof, ((size_t) dst->n_groups + src->n_groups + 1) * sizeof *gd);
	int32_t* const gp = realloc(dst->group_ptr, ((size_t) dst->n_groups + src->n_groups + 2) * sizeof *gp);
	double* const add = realloc(dst->add, ((size_t) old_adds + new_adds + 1) * sizeof *add);

	dst->group_dof = gd != NULL ? gd : dst->group_dof;
	dst->group_ptr = gp != NULL ? gp : dst->group_ptr;
	dst->add = add != NULL ? add : dst->add;

	if (gd == NULL || gp == NULL || add == NULL) {
		return -1;
	}

	for (int32_t g = 0; g < src->n_groups; g++) {
		dst->group_dof[dst->n_groups + g] = src->group_dof[g];
		dst->group_ptr[dst->n_groups + g] = old_adds + src->group_ptr[g];
	}

	memcpy(dst->add + old_adds, src->add, (size_t) new_adds * sizeof *add);

	dst->n_groups += src->n_groups;
	dst->group_ptr[dst->n_groups] = old_adds + new_adds;

	return 0;
}

int bfmx_job_create_batch(bfmx_job_t** out, bfm_sim_t** sims, size_t n_sims) {
	*out = NULL;

	if (n_sims == 0 || sims == NULL) {
		return -1;
	}

	bfm_state_t* const state = sims[0]->state;

	if (!bfmg_available()) {
		return BFMI_FAIL(state, "%s", bfmg_last_error());
	}

	if (bfmg_dist_world() > 1) {
		return BFMI_FAIL(state, "batches are independent systems: give each rank its own share instead of partitioning one batch");
	}

	bfmx_job_t* const job = calloc(1, sizeof *job);

	if (job == NULL) {
		return -1;
	}

	job->state = state;

	bc_op_t** sys_ops = NULL;
	size_t* sys_n_ops = NULL;
	size_t n_sys = 0;

	for (size_t i = 0; i < n_sims; i++) {
		if (sims[i]->kind == BFM_SIM_KIND_NONE) {
			continue; /* bfm_sim_run does nothing for these (sim.c:138-140) */
		}

		if (sims[i]->kind != BFM_SIM_KIND_PLANAR_STRAIN && sims[i]->kind != BFM_SIM_KIND_PLANAR_STRESS && sims[i]->kind != BFM_SIM_KIND_AXISYMMETRIC_STRAIN) {
			goto fail;
		}

		n_sys += sims[i]->n_instances;
	}

	if (n_sys == 0 || n_sys > INT32_MAX / 64) {
		BFMI_FAIL(state, "empty batch");
		goto fail;
	}

	job->sys = calloc(n_sys, sizeof *job->sys);
	job->h_tabs = calloc(n_sys, sizeof *job->h_tabs);
	job->ranges = calloc(n_sys, sizeof *job->ranges);
	job->status = calloc(n_sys, sizeof *job->status);
	sys_ops = calloc(n_sys, sizeof *sys_ops);
	sys_n_ops = calloc(n_sys, sizeof *sys_n_ops);

	if (job->sys == NULL || job->h_tabs == NULL || job->ranges == NULL || job->status == NULL || sys_ops == NULL || sys_n_ops == NULL) {
		goto fail;
	}

	/* layout */

	size_t nb_total = 0;
	size_t ne_total = 0;
	size_t max_forces = 0;
	size_t kind = 0;
	bool axisym = false;

	{
		size_t s = 0;

		for (size_t i = 0; i < n_sims; i++) {
			if (sims[i]->kind == BFM_SIM_KIND_NONE) {
				continue;
			}

			for (size_t k = 0; k < sims[i]->n_instances; k++, s++) {
				bfm_instance_t* const instance = sims[i]->instances[k];
				bfm_mesh_t const* const mesh = instance->obj->mesh;
				bool const axi = sims[i]->kind == BFM_SIM_KIND_AXISYMMETRIC_STRAIN;

				if (mesh->dim != 2 || (mesh->kind != BFM_ELEM_KIND_SIMPLEX && mesh->kind != BFM_ELEM_KIND_QUAD)) {
					goto fail; /* system.c:436-442 */
				}

				if (s == 0) {
					kind = mesh->kind;
					axisym = axi;
				}

				else if (mesh->kind != kind || axi != axisym) {
					BFMI_FAIL(state, "a batch must use one element kind and must not mix planar with axisymmetric problems");
					goto fail;
				}

				if ((int64_t) mesh->n_nodes > bfmg_batch_max_rows()) {
					BFMI_FAIL(state, "system %zu has %zu nodes; the one-CTA solver takes at most %d (run it with bfm_sim_run)", s, mesh->n_nodes, bfmg_batch_max_rows());
					goto fail;
				}

				job->sys[s].instance = instance;
				job->sys[s].mesh = mesh;
				job->sys[s].node_off = nb_total;

				job->ranges[s].row_lo = (int32_t) nb_total;
				job->ranges[s].row_hi = (int32_t) (nb_total + mesh->n_nodes);

				nb_total += (mesh->n_nodes + 31) / 32 * 32;
				ne_total += mesh->n_elems;
				max_forces = sims[i]->n_forces > max_forces ? sims[i]->n_forces : max_forces;
			}
		}
	}

	job->n_sys = (int32_t) n_sys;
	job->kind = axisym ? BFM_SIM_KIND_AXISYMMETRIC_STRAIN : BFM_SIM_KIND_PLANAR_STRESS;

	/* the combined mesh */

	bfm_mesh_t* const cm = &job->combined;

	cm->state = state;
	cm->dim = 2;
	cm->kind = kind;
	cm->n_nodes = nb_total;
	cm->n_elems = ne_total;
	cm->coords = calloc(nb_total * 2 + 1, sizeof *cm->coords);
	cm->elems = malloc((ne_total * kind + 1) * sizeof *cm->elems);

	if (cm->coords == NULL || cm->elems == NULL) {
		goto fail;
	}

	{
		size_t e_off = 0;

		for (size_t s = 0; s < n_sys; s++) {
			bfm_mesh_t const* const mesh = job->sys[s].mesh;
			size_t const off = job->sys[s].node_off;

			memcpy(cm->coords + 2 * off, mesh->coords, mesh->n_nodes * 2 * sizeof *cm->coords);

			for (size_t t = 0; t < mesh->n_elems * kind; t++) {
				if (mesh->elems[t] >= mesh->n_nodes) {
					BFMI_FAIL(state, "system %zu: connectivity points outside the node table", s);
					goto fail;
				}

				cm->elems[e_off * kind + t] = mesh->elems[t] + off;
			}

			e_off += mesh->n_elems;
		}
	}

	job->gmesh = cm;
	job->mesh = cm;

	double const t0 = now_ms();

	/* examples/benchmark.py-style loops run the same batch again and again: the symbolic plan of the combined mesh
	 * (and its device mirror) is kept as long as the combined connectivity is the same - 36 ms -> a few ms per call
	 * at 1024 x 335 nodes.  The combined mesh object itself dies with the job, so the key is its content. */

	static bfmi_plan_t* batch_plan;
	static uint64_t batch_hash;

	uint64_t const hash = bfmi_mesh_hash(cm);

	if (batch_plan != NULL && batch_hash == hash && (size_t) batch_plan->nb == nb_total && batch_plan->n_elems == ne_total && batch_plan->kind == (int) kind) {
		job->plan = batch_plan;
		bfmi_plan_retain(job->plan);
	}

	else {
		job->plan = bfmi_plan_for_mesh(state, cm);

		if (job->plan == NULL || bfmi_plan_upload(state, job->plan) < 0) {
			goto fail;
		}

		job->stats.ms_plan = (float) (now_ms() - t0);
		job->stats.h2d_bytes += job->plan->h2d_bytes;
		job->stats.d2h_bytes += job->plan->d2h_bytes;

		bfmi_plan_release(batch_plan);
		batch_plan = job->plan;
		batch_hash = hash;
		bfmi_plan_retain(batch_plan);
	}

	job->pat = job->plan->dev;

	job->h_slice_tab = calloc((size_t) job->plan->n_slices + 1, sizeof *job->h_slice_tab);

	if (job->h_slice_tab == NULL) {
		goto fail;
	}

	/* per-system tables and work lists */

	size_t max_ops = 0;

	{
		size_t s = 0;

		for (size_t i = 0; i < n_sims; i++) {
			if (sims[i]->kind == BFM_SIM_KIND_NONE) {
				continue;
			}

			for (size_t k = 0; k < sims[i]->n_instances; k++, s++) {
				bfm_mesh_t const* const mesh = job->sys[s].mesh;
				size_t const off = job->sys[s].node_off;

				if (fill_tables(state, sims[i]->kind, job->sys[s].instance, mesh, sims[i]->n_forces, sims[i]->forces, &job->h_tabs[s], &job->h_nforce, max_forces, nb_total, off) < 0) {
					goto fail;
				}

				for (size_t slice = off / 32; slice < (off + mesh->n_nodes + 31) / 32; slice++) {
					job->h_slice_tab[slice] = (int32_t) s;
				}

				if (build_ops_for(mesh, sims[i]->kind, job->sys[s].instance, &sys_ops[s], &sys_n_ops[s]) < 0) {
					goto fail;
				}

				for (size_t c = 0; c < sys_n_ops[s]; c++) {
					shift_op(&sys_ops[s][c], off);
				}

				max_ops = sys_n_ops[s] > max_ops ? sys_n_ops[s] : max_ops;
			}
		}
	}

	job->tab = job->h_tabs[0]; /* kind / axisym for the launcher */

	job->ops = calloc(2 * max_ops + 1, sizeof *job->ops);

	if (job->ops == NULL) {
		goto fail;
	}

	for (size_t c = 0; c < max_ops; c++) {
		bc_op_t* const dir = &job->ops[job->n_ops];
		bc_op_t* const add = &job->ops[job->n_ops + 1];

		dir->kind = OP_DIRICHLET;
		add->kind = OP_ADD;
		job->n_ops += 2; /* both are released with the job whatever happens below */

		for (size_t s = 0; s < n_sys; s++) {
			if (c < sys_n_ops[s] && append_op(sys_ops[s][c].kind == OP_DIRICHLET ? dir : add, &sys_ops[s][c]) < 0) {
				goto fail;
			}
		}

		if (affected_rows(job->plan, dir) < 0) {
			goto fail;
		}
	}

	for (size_t s = 0; s < n_sys; s++) {
		free_ops(sys_ops[s], sys_n_ops[s]);
	}

	free(sys_ops);
	free(sys_n_ops);
	sys_ops = NULL;

	job->table_forces = max_forces;

	if (
		job_alloc_device(job, max_forces) < 0 ||
		bfmg_alloc((void**) &job->d_tabs, n_sys * sizeof *job->h_tabs) < 0 ||
		bfmg_alloc((void**) &job->d_slice_tab, ((size_t) job->plan->n_slices + 1) * sizeof(int32_t)) < 0
	) {
		goto fail;
	}

	job->stats.n_dofs = 0;

	for (size_t s = 0; s < n_sys; s++) {
		job->stats.n_dofs += 2 * job->sys[s].mesh->n_nodes;
	}

	job->stats.n_dofs_owned = job->stats.n_dofs;
	job->stats.n_ranks = 1;
	job->stats.n_blocks = (size_t) job->plan->n_blocks;
	job->stats.n_slots = (size_t) job->plan->n_slots;

	*out = job;
	return 0;

fail:

	if (sys_ops != NULL) {
		for (size_t s = 0; s < n_sys; s++) {
			free_ops(sys_ops[s], sys_n_ops != NULL ? sys_n_ops[s] : 0);
		}
	}

	free(sys_ops);
	free(sys_n_ops);

	bfmx_job_destroy(job);
	return -1;
}

static int batch_download(bfmx_job_t* job) {
	size_t const bytes = (size_t) job->plan->nb * 2 * sizeof(double);
	double* const staged = malloc(bytes + 8);

	if (staged == NULL) {
		return -1;
	}

	int const t0 = bfmg_tick();

	if (bfmg_download(staged, job->d_x, bytes) < 0) {
		free(staged);
		return BFMI_FAIL(job->state, "download failed: %s", bfmg_last_error());
	}

	int const t1 = bfmg_tick();

	for (int32_t s = 0; s < job->n_sys; s++) { /* sim.c:127-131, per system */
		memcpy(job->sys[s].instance->effects, staged + 2 * job->sys[s].node_off, job->sys[s].mesh->n_nodes * 2 * sizeof(double));
	}

	free(staged);

	job->stats.ms_download = bfmg_lap(t0, t1);
	job->stats.d2h_bytes += bytes;

	return 0;
}

int bfmx_job_batch_size(bfmx_job_t* job) {
	return job->n_sys;
}

int bfmx_job_batch_status(bfmx_job_t* job, size_t system, bfmx_batch_status_t* out) {
	if (system >= (size_t) job->n_sys || !job->solved) {
		return -1;
	}

	out->iterations = job->status[system].iterations;
	out->converged = job->status[system].converged;
	out->rel_residual = job->status[system].rel_residual;
	out->true_rel_residual = job->status[system].true_rel_residual;
	out->backward_error = job->status[system].backward_error;

	return 0;
}

/* what examples/benchmark.py's loop does for many simulations at once */
int bfmx_sim_run_batch(bfm_sim_t** sims, size_t n_sims) {
	bfmx_job_t* job;

	if (bfmx_job_create_batch(&job, sims, n_sims) < 0) {
		return -1;
	}

	int const rv = bfmx_job_upload(job) < 0 || bfmx_job_assemble(job) < 0 || bfmx_job_solve(job) < 0 || bfmx_job_download(job) < 0 ? -1 : 0;

	bfmx_publish_stats(&job->stats);
	bfmx_job_destroy(job);

	return rv;
}

int bfmx_job_destroy(bfmx_job_t* job) {
	if (job == NULL) {
		return 0;
	}

	free_ops(job->ops, job->n_ops);
	free(job->h_nforce);

	bfmg_free(job->d_coords);
	bfmg_free(job->d_nforce);
	bfmg_free(job->d_val);
	bfmg_free(job->d_b);
	bfmg_free(job->d_x);
	bfmg_free(job->d_stamp);
	bfmg_free(job->d_cval);
	bfmg_free(job->d_xg);
	bfmg_free(job->d_xp);
	bfmi_coarse_release(job->coarse);
	bfmi_hier_release(job->hier);
	bfmg_free(job->d_tabs);
	bfmg_free(job->d_slice_tab);

	if (job->n_sys > 0) {
		bfmi_plan_forget(&job->combined);
	}

	free(job->combined.coords);
	free(job->combined.elems);
	free(job->sys);
	free(job->h_tabs);
	free(job->h_slice_tab);
	free(job->ranges);
	free(job->status);

	bfmi_plan_release(job->plan);
	bfmi_part_release(job->part);
	bfmi_renum_release(job->renum);
	free(job);

	return 0;
}

#define UP(dst, src, count, type) (bytes += (size_t) (count) * sizeof(type), bfmg_upload((dst), (src), (size_t) (count) * sizeof(type)))

int bfmx_job_upload(bfmx_job_t* job) {
	size_t const nb = (size_t) job->plan->nb;
	size_t bytes = 0;
	int const t0 = bfmg_tick();
	int rv = 0;

	if (nb * 2 * sizeof(double) >= ((size_t) 32 << 20) && job->n_sys == 0) {
		bfmg_host_pin(job->mesh->coords, nb * 2 * sizeof(double)); /* lives as long as the mesh: pinned once, unpinned by its destroy */
	}

	rv |= UP(job->d_coords, job->mesh->coords, nb * 2, double);

	if (job->h_nforce != NULL) {
		rv |= UP(job->d_nforce, job->h_nforce, job->table_forces * nb * 2, double);
	}

	for (size_t i = 0; i < job->n_ops; i++) {
		bc_op_t* const op = &job->ops[i];

		if (op->kind == OP_DIRICHLET) {
			rv |= UP(op->d_dofs, op->dofs, op->n_dofs, int32_t);
			rv |= UP(op->d_vals, op->vals, op->n_dofs, double);
			rv |= UP(op->d_rows, op->rows, op->n_rows, int32_t);
		}

		else if (op->n_groups > 0) {
			rv |= UP(op->d_group_dof, op->group_dof, op->n_groups, int32_t);
			rv |= UP(op->d_group_ptr, op->group_ptr, op->n_groups + 1, int32_t);
			rv |= UP(op->d_add, op->add, op->group_ptr[op->n_groups], double);
		}
	}

	if (job->n_sys > 0) {
		rv |= UP(job->d_tabs, job->h_tabs, job->n_sys, bfmg_asm_tables_t);
		rv |= UP(job->d_slice_tab, job->h_slice_tab, job->plan->n_slices, int32_t);
	}

	if (job->d_stamp != NULL) {
		rv |= bfmg_zero(job->d_stamp, nb * 2 * sizeof(int32_t));
	}

	int const t1 = bfmg_tick();

	if (rv < 0 || bfmg_sync() < 0) {
		return BFMI_FAIL(job->state, "upload failed: %s", bfmg_last_error());
	}

	job->stats.ms_upload = bfmg_lap(t0, t1);
	job->stats.h2d_bytes += bytes;
	job->uploaded = true;

	return 0;
}

int bfmx_job_assemble(bfmx_job_t* job) {
	if (!job->uploaded) {
		return -1;
	}

	size_t const before = bfmg_launch_count();
	int const t0 = bfmg_tick();

	if (bfmg_assemble(&job->plan->dev, &job->tab, job->d_coords, job->d_nforce, job->d_val, job->d_b, job->d_tabs, job->d_slice_tab) < 0) {
		return BFMI_FAIL(job->state, "assembly failed: %s", bfmg_last_error());
	}

	int const t1 = bfmg_tick();

	for (size_t i = 0; i < job->n_ops; i++) {
		bc_op_t const* const op = &job->ops[i];
		int rv;

		if (op->kind == OP_DIRICHLET) {
			rv = bfmg_bc_dirichlet(&job->plan->dev, job->d_val, job->d_b, job->d_stamp, job->d_cval, (int32_t) i + 1, op->d_dofs, op->d_vals, op->n_dofs, op->d_rows, op->n_rows);
		}

		else {
			rv = bfmg_bc_add(job->d_b, op->d_group_dof, op->d_group_ptr, op->d_add, op->n_groups);
		}

		if (rv < 0) {
			return BFMI_FAIL(job->state, "boundary conditions failed: %s", bfmg_last_error());
		}
	}

	int const t2 = bfmg_tick();

	if (bfmg_sync() < 0) {
		return BFMI_FAIL(job->state, "assembly failed: %s", bfmg_last_error());
	}

	job->stats.ms_assemble = bfmg_lap(t0, t1);
	job->stats.ms_bc = bfmg_lap(t1, t2);
	job->stats.kernel_launches += bfmg_launch_count() - before;
	job->assembled = true;

	return 0;
}

int bfmx_job_solve(bfmx_job_t* job) {
	if (!job->assembled) {
		return -1;
	}

	bfmg_pcg_opts_t opts;
	bfmg_pcg_result_t res = {0};

	bfmi_pcg_options(job->stats.n_dofs, &opts);

	if (job->n_sys > 0) {
		size_t const before = bfmg_launch_count();
		float ms = 0;

		if (bfmg_pcg_batch(&job->pat, job->d_val, job->d_b, job->d_x, &opts, job->n_sys, job->ranges, job->status, &ms) < 0) {
			return BFMI_FAIL(job->state, "batched PCG failed: %s", bfmg_last_error());
		}

		/* the job's figures are the worst over the batch; per-system ones: bfmx_job_batch_status */

		job->stats.cg_iterations = 0;
		job->stats.cg_converged = 1;
		job->stats.cg_rel_residual = job->stats.cg_true_rel_residual = job->stats.cg_backward_error = 0;

		int32_t failed = 0;

		for (int32_t i = 0; i < job->n_sys; i++) {
			bfmg_batch_status_t* const st = &job->status[i];

			if (st->converged == 1 && st->backward_error > opts.true_tol) {
				st->converged = 0;
			}

			failed += st->converged != 1;

			job->stats.cg_iterations = st->iterations > job->stats.cg_iterations ? st->iterations : job->stats.cg_iterations;
			job->stats.cg_converged = st->converged < job->stats.cg_converged ? st->converged : job->stats.cg_converged;
			job->stats.cg_rel_residual = fmax(job->stats.cg_rel_residual, st->rel_residual);
			job->stats.cg_true_rel_residual = fmax(job->stats.cg_true_rel_residual, st->true_rel_residual);
			job->stats.cg_backward_error = fmax(job->stats.cg_backward_error, st->backward_error);
		}

		job->stats.cg_restarts = 0;
		job->stats.ms_solve = ms;
		job->stats.kernel_launches += bfmg_launch_count() - before;
		job->solved = true;

		if (failed > 0) {
			return BFMI_FAIL(job->state, "%d of the %d systems of the batch did not converge", failed, job->n_sys);
		}

		return 0;
	}

	if (takes_one_cta(job)) {
		bfmg_batch_range_t const range = {0, job->plan->nb};
		bfmg_batch_status_t st;
		size_t const before = bfmg_launch_count();

		if (bfmg_pcg_batch(&job->pat, job->d_val, job->d_b, job->d_x, &opts, 1, &range, &st, &res.ms) < 0) {
			return BFMI_FAIL(job->state, "PCG failed: %s", bfmg_last_error());
		}

		res.iterations = st.iterations;
		res.restarts = 0;
		res.rel_residual = st.rel_residual;
		res.true_rel_residual = st.true_rel_residual;
		res.backward_error = st.backward_error;
		res.converged = st.converged == 1 && st.backward_error > opts.true_tol ? 0 : st.converged;
		res.launches = bfmg_launch_count() - before;
	}

	else if (bfmg_pcg(&job->pat, job->d_val, job->d_b, job->d_x, &opts, &res, job->part != NULL ? &job->halo : NULL, job->coarse != NULL ? &job->coarse->dev : NULL, job->hier != NULL ? &job->hier->dev : NULL) < 0) {
		return BFMI_FAIL(job->state, "PCG failed: %s", bfmg_last_error());
	}

	job->stats.cg_iterations = res.iterations;
	job->stats.cg_restarts = res.restarts;
	job->stats.cg_converged = res.converged;
	job->stats.cg_rel_residual = res.rel_residual;
	job->stats.cg_true_rel_residual = res.true_rel_residual;
	job->stats.cg_backward_error = res.backward_error;
	job->stats.ms_solve = res.ms;
	job->stats.ms_solve_setup = res.ms_setup;
	job->stats.coarse_dim = (size_t) res.coarse_dim;
	job->stats.mg_levels = res.mg_levels;
	job->stats.uses_peer_memory = res.peer_memory;
	job->stats.kernel_launches += res.launches;
	job->solved = true;

	if (res.converged != 1) {
		return BFMI_FAIL(job->state, "PCG stopped after %d iterations without converging (relative residual %.3e, recomputed %.3e, backward error %.3e)", res.iterations, res.rel_residual, res.true_rel_residual, res.backward_error);
	}

	return 0;
}

int bfmx_job_download(bfmx_job_t* job) {
	if (!job->solved) {
		return -1;
	}

	if (job->n_sys > 0) {
		return batch_download(job);
	}

	size_t const bytes = job->stats.n_dofs * sizeof(double);
	int const t0 = bfmg_tick();
	double const* d_src = job->d_x;

	/* several GPUs: every rank receives every owned block, so that each process ends with the complete
	 * field, exactly as a single-process caller of the reference does */

	if (job->part != NULL) {
		size_t first[BFMG_DIST_MAX_RANKS + 1];

		for (int r = 0; r <= job->part->world; r++) {
			first[r] = bfmi_part_first_node(job->part->n_nodes, job->part->world, r);
		}

		if (bfmg_dist_gather_blocks(job->d_x + 2 * (size_t) job->part->own_begin, first, job->d_xg) < 0) {
			return BFMI_FAIL(job->state, "gathering the displacements failed: %s", bfmg_last_error());
		}

		d_src = job->d_xg;
	}

	/* an internally renumbered job: node a of the caller is node to_new[a] here */

	if (job->renum != NULL) {
		if (bfmg_gather_blocks(job->d_xp, d_src, job->renum->d_to_new, job->renum->n_nodes) < 0) {
			return BFMI_FAIL(job->state, "reordering the displacements failed: %s", bfmg_last_error());
		}

		d_src = job->d_xp;
	}

	/* effects[node * dim + k] = x[node * dim + k] (sim.c:127-131): same interleaving, one copy */

	if (bytes >= ((size_t) 32 << 20)) {
		bfmg_host_pin(job->instance->effects, bytes); /* lives as long as the instance: pinned once, unpinned by its destroy */
	}

	if (bfmg_download(job->instance->effects, d_src, bytes) < 0) {
		return BFMI_FAIL(job->state, "download failed: %s", bfmg_last_error());
	}

	int const t1 = bfmg_tick();

	job->stats.ms_download = bfmg_lap(t0, t1);
	job->stats.d2h_bytes += bytes;

	return 0;
}

int bfmx_job_stats(bfmx_job_t* job, bfmx_stats_t* out) {
	*out = job->stats;
	return 0;
}

int bfmx_job_spmv_time(bfmx_job_t* job, int reps, float* ms_per_launch) {
	if (!job->assembled) {
		return -1;
	}

	return bfmg_spmv_time(&job->pat, job->d_val, reps, ms_per_launch);
}

int bfmx_job_read(bfmx_job_t* job, double* b, double* x) {
	size_t const bytes = (size_t) job->plan->nb * 2 * sizeof(double); /* local rows on a partitioned job */

	if (job->renum != NULL && job->part == NULL) { /* internally renumbered: hand both back in the caller's numbering */
		if (b != NULL && (!job->assembled || bfmg_gather_blocks(job->d_xp, job->d_b, job->renum->d_to_new, job->renum->n_nodes) < 0 || bfmg_download(b, job->d_xp, bytes) < 0)) {
			return -1;
		}

		if (x != NULL && (!job->solved || bfmg_gather_blocks(job->d_xp, job->d_x, job->renum->d_to_new, job->renum->n_nodes) < 0 || bfmg_download(x, job->d_xp, bytes) < 0)) {
			return -1;
		}

		return 0;
	}

	if (b != NULL && (!job->assembled || bfmg_download(b, job->d_b, bytes) < 0)) {
		return -1;
	}

	if (x != NULL && (!job->solved || bfmg_download(x, job->d_x, bytes) < 0)) {
		return -1;
	}

	return 0;
}

/* ---- bfm_sim_run (reference sim.c:137-155) ---------------------------------------------------------- */

int bfm_sim_run(bfm_sim_t* sim) {
	if (sim->kind == BFM_SIM_KIND_NONE) {
		return 0;
	}

	if (sim->kind != BFM_SIM_KIND_PLANAR_STRAIN && sim->kind != BFM_SIM_KIND_PLANAR_STRESS && sim->kind != BFM_SIM_KIND_AXISYMMETRIC_STRAIN) {
		return -1;
	}

	bool const verbose = getenv("BFM_JOB_VERBOSE") != NULL;

	for (size_t i = 0; i < sim->n_instances; i++) {
		bfmx_job_t* job;
		double t[7];

		t[0] = now_ms();

		if (bfmx_job_create(&job, sim, i) < 0) {
			return -1;
		}

		t[1] = now_ms();

		int rv = bfmx_job_upload(job) < 0 ? -1 : 0;

		t[2] = now_ms();

		rv = rv < 0 || bfmx_job_assemble(job) < 0 ? -1 : 0;

		t[3] = t[4] = now_ms();

		if (rv == 0) {
			/* the reference ignores bfm_matrix_solve's status and always fills instance->effects (sim.c:122-131).  A PCG that
			 * misses its tolerance is reported as -1 here (state->err says why) - but the best iterate is still delivered,
			 * on every rank alike, so callers that ignore the status get what the reference would have given them */
			int const solved = bfmx_job_solve(job);

			t[4] = now_ms();

			if (job->solved && bfmx_job_download(job) < 0) {
				rv = -1;
			}

			rv = solved < 0 ? -1 : rv;
		}

		t[5] = now_ms();

		bfmx_publish_stats(&job->stats);
		bfmx_job_destroy(job);

		t[6] = now_ms();

		if (verbose) {
			fprintf(stderr, "[sim_run] create %.1f  upload %.1f  assemble %.1f  solve %.1f  download %.1f  destroy %.1f ms (host wall clock)\n", t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]);
		}

		if (rv < 0) {
			return -1;
		}
	}

	return 0;
}

/* ---- bfm_system_* (reference system.c) ---------------------------------------------------------------- */

int bfm_system_create(bfm_system_t* system, bfm_state_t* state, size_t n) { /* system.c:5-34 */
	system->state = state;
	system->n = n;

	if (bfm_perm_create(&system->perm, state, n) < 0) {
		return -1;
	}

	if (bfm_matrix_full_create(&system->A, state, BFM_MATRIX_MAJOR_ROW, n) < 0) {
		bfm_perm_destroy(&system->perm);
		return -1;
	}

	if (bfm_vec_create(&system->b, state, n) < 0) {
		bfm_matrix_destroy(&system->A);
		bfm_perm_destroy(&system->perm);
		return -1;
	}

	return 0;
}

int bfm_system_destroy(bfm_system_t* system) { /* system.c:36-42 */
	bfm_perm_destroy(&system->perm);
	bfm_matrix_destroy(&system->A);
	bfm_vec_destroy(&system->b);

	return 0;
}

/* assembly + BCs on the GPU; A is handed back as a CSR matrix that keeps the device buffer */
static int system_create_gpu(bfm_system_t* system, bfm_sim_kind_t kind, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces) {
	bfm_state_t* const state = instance->state;
	bfmx_job_t* job;

	if (job_create(&job, state, kind, instance, n_forces, forces, false) < 0) { /* the staged API is single-GPU */
		return -1;
	}

	size_t const n = job->stats.n_dofs;
	int rv = -1;
	bool have_perm = false, have_b = false;

	/* a failed call leaves a zeroed system behind (bfm_system_destroy on it is harmless), never a half-built one */
	memset(system, 0, sizeof *system);

	system->state = state;
	system->n = n;

	if (bfmx_job_upload(job) < 0 || bfmx_job_assemble(job) < 0) {
		goto done;
	}

	if (bfm_perm_create(&system->perm, state, n) < 0) {
		goto done;
	}

	have_perm = true;

	if (bfm_vec_create(&system->b, state, n) < 0) {
		goto done;
	}

	have_b = true;

	if (bfmg_download(system->b.data, job->d_b, n * sizeof(double)) < 0 || bfmi_csr_wrap(&system->A, state, job->plan, job->d_val) < 0) {
		goto done;
	}

	job->d_val = NULL; /* now owned by the matrix */
	job->stats.d2h_bytes += n * sizeof(double);
	rv = 0;

done:

	if (rv < 0) {
		if (have_b) {
			bfm_vec_destroy(&system->b);
		}

		if (have_perm) {
			bfm_perm_destroy(&system->perm);
		}

		memset(system, 0, sizeof *system);
		system->state = state;
	}

	bfmx_publish_stats(&job->stats);
	bfmx_job_destroy(job);

	return rv;
}

int bfm_system_create_planar_strain(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces) {
	return system_create_gpu(system, BFM_SIM_KIND_PLANAR_STRAIN, instance, n_forces, forces);
}

int bfm_system_create_planar_stress(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces) {
	return system_create_gpu(system, BFM_SIM_KIND_PLANAR_STRESS, instance, n_forces, forces);
}

int bfm_system_create_axisymmetric_strain(bfm_system_t* system, bfm_instance_t* instance, size_t n_forces, bfm_force_t** forces) {
	return system_create_gpu(system, BFM_SIM_KIND_AXISYMMETRIC_STRAIN, instance, n_forces, forces);
}

/* reference system.c:44-81 */
int bfm_system_renumber(bfm_system_t* system) {
	if (bfm_perm_rcm(&system->perm, &system->A) < 0) {
		return -1;
	}

	if (bfm_perm_perm_matrix(&system->perm, &system->A, false) < 0 || bfm_perm_perm_vec(&system->perm, &system->b, false) < 0) {
		return -1;
	}

	if (system->A.kind != BFM_MATRIX_KIND_FULL) {
		return 0; /* CSR: renumbered logically, nothing to convert */
	}

	/* dense -> band, as the reference does */

	bfm_matrix_t band;

	if (bfm_matrix_band_create(&band, system->state, system->A.major, system->A.m, bfm_matrix_bandwidth(&system->A)) < 0) {
		return -1;
	}

	if (bfm_matrix_copy(&band, &system->A) < 0) {
		bfm_matrix_destroy(&band);
		return -1;
	}

	bfm_matrix_destroy(&system->A);
	system->A = band;

	return 0;
}
