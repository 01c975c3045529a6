/*
 * core.c - the plain-host part of libbfm's API: state, vectors, materials, shapes, rules, objects,
 * conditions, instances, forces and the simulation container.
 *
 * None of this is on the GPU hot path; it is the struct plumbing either side of it (SURVEY.md
 * section 2, rows marked "input").  Behaviour follows the reference function by function - the
 * citation next to each one names the lines it replaces - including the quirks callers may
 * depend on (they are called out where they occur).
 */
#include "internal.h"

#include <stdio.h>
#include <string.h>

/* ---- state (reference state.c:6-46) ---------------------------------------------------------- */

int bfm_state_create(bfm_state_t* state) {
	memset(state, 0, sizeof *state);

	state->alloc = malloc;
	state->realloc = realloc;
	state->free = free;

	return 0;
}

int bfm_state_destroy(bfm_state_t* state) {
	(void) state;
	return 0;
}

int bfm_set_alloc(bfm_state_t* state, bfm_alloc_t alloc) {
	state->alloc = alloc;
	return 0;
}

int bfm_set_realloc(bfm_state_t* state, bfm_realloc_t realloc) {
	state->realloc = realloc;
	return 0;
}

int bfm_set_free(bfm_state_t* state, bfm_free_t free) {
	state->free = free;
	return 0;
}

int bfm_err_print(bfm_state_t* state) {
	bfm_err_t const* const err = &state->err;

	if (err->has) {
		printf("[BFM %s:%zu (%s)] %s\n", err->file, err->line, err->func, err->msg);
	}

	return 0;
}

/* The reference declares bfm_err_t but never writes it; we use it to say why the GPU path failed.
 * msg points into a static buffer so that nothing has to be freed through the state allocator. */
int bfmi_fail(bfm_state_t* state, char const* file, char const* func, size_t line, char const* fmt, ...) {
	static char buf[512];
	va_list ap;

	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);

	if (state != NULL) {
		state->err.has = true;
		state->err.msg = buf;
		state->err.file = file;
		state->err.func = func;
		state->err.line = line;
	}

	if (getenv("BFM_QUIET") == NULL) {
		fprintf(stderr, "[BFM %s:%zu (%s)] %s\n", file, line, func, buf);
	}

	return -1;
}

/* ---- vectors (reference vec.c:5-32) ---------------------------------------------------------- */

int bfm_vec_create(bfm_vec_t* vec, bfm_state_t* state, size_t n) {
	vec->state = state;
	vec->n = n;
	vec->data = state->alloc(n * sizeof *vec->data);

	if (vec->data == NULL) {
		return -1;
	}

	memset(vec->data, 0, n * sizeof *vec->data);
	return 0;
}

int bfm_vec_copy(bfm_vec_t* vec, bfm_vec_t* src) {
	memcpy(vec->data, src->data, src->n * sizeof *src->data);
	return 0;
}

int bfm_vec_destroy(bfm_vec_t* vec) {
	vec->state->free(vec->data);
	return 0;
}

/* ---- materials (reference material.c:5-40) --------------------------------------------------- */

int bfm_material_create(bfm_material_t* material, bfm_state_t* state, char* name, double rho, double E, double nu) {
	memset(material, 0, sizeof *material);

	material->state = state;
	material->name = strdup(name); /* released with state->free, as in the reference (material.c:12,28) */

	if (material->name == NULL) {
		return -1;
	}

	material->rho = rho;
	material->E = E;
	material->nu = nu;

	return 0;
}

int bfm_material_destroy(bfm_material_t* material) {
	material->state->free(material->name);
	return 0;
}

int bfm_material_set_colour(bfm_material_t* material, double r, double g, double b, double a) {
	material->colour = (bfm_colour_t) {.r = r, .g = g, .b = b, .a = a};
	return 0;
}

/* ---- shape functions (reference shape.c:3-105) ------------------------------------------------
 * Node order of the quad is (+,+), (-,+), (-,-), (+,-).  The arithmetic (including the "/ 4" after
 * the product) is kept exactly: the GPU assembly consumes tables produced by these functions and
 * the stiffness values must round like the reference's. */

static int phi_default(bfm_shape_t* shape, double* point, double* phi) {
	if (shape->dim != 2) {
		return -1;
	}

	double const s = point[0];
	double const t = point[1];

	switch (shape->kind) {
	case BFM_ELEM_KIND_SIMPLEX:
		phi[0] = 1 - s - t;
		phi[1] = s;
		phi[2] = t;
		return 0;

	case BFM_ELEM_KIND_QUAD:
		phi[0] = (1 + s) * (1 + t) / 4;
		phi[1] = (1 - s) * (1 + t) / 4;
		phi[2] = (1 - s) * (1 - t) / 4;
		phi[3] = (1 + s) * (1 - t) / 4;
		return 0;

	case BFM_ELEM_KIND_QUADRATIC_TRIANGLE:
		/* the reference fills the six P2 values and then still reports failure (shape.c:29-37) */
		phi[0] = 1 - 3 * (s + t) + 2 * (s + t) * (s + t);
		phi[1] = s * (2 * s - 1);
		phi[2] = t * (2 * t - 1);
		phi[3] = 4 * s * (1 - s - t);
		phi[4] = 4 * s * t;
		phi[5] = 4 * t * (1 - s - t);
		return -1;
	}

	return -1;
}

static int dphi_default(bfm_shape_t* shape, size_t wrt, double* point, double* d) {
	if (shape->dim != 2) {
		return -1;
	}

	double const s = point[0];
	double const t = point[1];
	bool const dxsi = wrt == 0;

	switch (shape->kind) {
	case BFM_ELEM_KIND_SIMPLEX:
		d[0] = -1;
		d[1] = dxsi ? 1 : 0;
		d[2] = dxsi ? 0 : 1;
		return 0;

	case BFM_ELEM_KIND_QUAD:
		d[0] = (dxsi ? 1 + t : 1 + s) / 4;
		d[1] = (dxsi ? -1 - t : 1 - s) / 4;
		d[2] = (dxsi ? -1 + t : -1 + s) / 4;
		d[3] = (dxsi ? 1 - t : -1 - s) / 4;
		return 0;

	case BFM_ELEM_KIND_QUADRATIC_TRIANGLE:
		d[0] = -3 + 4 * (s + t);
		d[1] = dxsi ? 4 * s - 1 : 0;
		d[2] = dxsi ? 0 : 4 * t - 1;
		d[3] = dxsi ? 4 - 8 * s - 4 * t : -4 * s;
		d[4] = dxsi ? 4 * t : 4 * s;
		d[5] = dxsi ? -4 * t : 4 - 4 * s - 8 * t;
		return -1; /* shape.c:66-84: same "filled but failed" behaviour */
	}

	return -1;
}

int bfm_shape_create(bfm_shape_t* shape, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind) {
	shape->state = state;
	shape->dim = dim;
	shape->kind = kind;
	shape->phi = phi_default;
	shape->dphi = dphi_default;

	return 0;
}

int bfm_shape_destroy(bfm_shape_t* shape) {
	(void) shape;
	return 0;
}

/* ---- integration rules (reference rule.c:5-151) ---------------------------------------------- */

int bfm_rule_create(bfm_rule_t* rule, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind, size_t n_points) {
	memset(rule, 0, sizeof *rule);

	rule->state = state;
	rule->dim = dim;
	rule->kind = kind;
	rule->n_points = n_points;

	rule->weights = state->alloc(n_points * sizeof *rule->weights);
	rule->points = state->alloc(n_points * sizeof *rule->points);

	if (rule->weights == NULL || rule->points == NULL) {
		goto fail;
	}

	memset(rule->weights, 0, n_points * sizeof *rule->weights);
	memset(rule->points, 0, n_points * sizeof *rule->points);

	for (size_t i = 0; i < n_points; i++) {
		rule->points[i] = state->alloc(dim * sizeof **rule->points);

		if (rule->points[i] == NULL) {
			goto fail;
		}

		memset(rule->points[i], 0, dim * sizeof **rule->points);
	}

	if (bfm_shape_create(&rule->shape, state, dim, kind) < 0) {
		goto fail;
	}

	return 0;

fail:

	for (size_t i = 0; rule->points != NULL && i < n_points; i++) {
		if (rule->points[i] != NULL) {
			state->free(rule->points[i]);
		}
	}

	if (rule->points != NULL) {
		state->free(rule->points);
	}

	if (rule->weights != NULL) {
		state->free(rule->weights);
	}

	return -1;
}

int bfm_rule_destroy(bfm_rule_t* rule) {
	bfm_state_t* const state = rule->state;

	/* (the reference releases only the first `dim` points, rule.c:76-78, and leaks the others) */

	for (size_t i = 0; i < rule->n_points; i++) {
		state->free(rule->points[i]);
	}

	state->free(rule->points);
	state->free(rule->weights);

	return bfm_shape_destroy(&rule->shape);
}

int bfm_rule_create_gauss_legendre(bfm_rule_t* rule, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind) {
	if (dim != 2 || (kind != BFM_ELEM_KIND_SIMPLEX && kind != BFM_ELEM_KIND_QUAD)) {
		return -1;
	}

	if (bfm_rule_create(rule, state, dim, kind, (size_t) kind) < 0) { /* n_points == kind, rule.c:126 */
		return -1;
	}

	if (kind == BFM_ELEM_KIND_SIMPLEX) {
		/* 3-point rule: (1/6,1/6), (2/3,1/6), (1/6,2/3), weight 1/6 each; 2/3 is formed as 1 - 1/3
		 * (rule.c:85-102) */

		double const sixth = 1. / 6;
		double const two_thirds = 1 - 1. / 3;

		for (size_t g = 0; g < 3; g++) {
			rule->weights[g] = sixth;
			rule->points[g][0] = g == 1 ? two_thirds : sixth;
			rule->points[g][1] = g == 2 ? two_thirds : sixth;
		}
	}

	else {
		/* 2x2 rule at +-0.577350269189626 - the reference's 15-digit literal, NOT sqrt(1/3)
		 * (rule.c:92); order (-,+), (-,-), (+,-), (+,+), unit weights (rule.c:104-111) */

		double const s = 0.577350269189626;
		double const sx[4] = {-s, -s, s, s};
		double const sy[4] = {s, -s, -s, s};

		for (size_t g = 0; g < 4; g++) {
			rule->weights[g] = 1;
			rule->points[g][0] = sx[g];
			rule->points[g][1] = sy[g];
		}
	}

	return 0;
}

/* ---- objects (reference obj.c:3-16) ---------------------------------------------------------- */

int bfm_obj_create(bfm_obj_t* obj, bfm_state_t* state, bfm_mesh_t* mesh, bfm_material_t* material, bfm_rule_t* rule) {
	*obj = (bfm_obj_t) {.state = state, .mesh = mesh, .material = material, .rule = rule};
	return 0;
}

int bfm_obj_destroy(bfm_obj_t* obj) {
	(void) obj;
	return 0;
}

/* ---- conditions (reference condition.c:5-29) ------------------------------------------------- */

int bfm_condition_create(bfm_condition_t* condition, bfm_state_t* state, bfm_mesh_t* mesh, bfm_condition_kind_t kind) {
	/* `value` is deliberately left alone: the reference never initialises it and pybfm relies on the
	 * zero it gets from ffi.new (pybfm/bfm/condition.py:18-19) */

	condition->state = state;
	condition->mesh = mesh;
	condition->kind = kind;
	condition->nodes = state->alloc(mesh->n_nodes * sizeof *condition->nodes);

	if (condition->nodes == NULL) {
		return -1;
	}

	memset(condition->nodes, 0, mesh->n_nodes * sizeof *condition->nodes);
	return 0;
}

int bfm_condition_destroy(bfm_condition_t* condition) {
	condition->state->free(condition->nodes);
	return 0;
}

/* ---- instances (reference instance.c:6-73) --------------------------------------------------- */

int bfm_instance_create(bfm_instance_t* instance, bfm_state_t* state, bfm_obj_t* obj) {
	memset(instance, 0, sizeof *instance);

	instance->state = state;
	instance->obj = obj;
	instance->n_effects = obj->mesh->n_nodes * obj->mesh->dim;
	instance->effects = state->alloc(instance->n_effects * sizeof *instance->effects);

	if (instance->effects == NULL) {
		return -1;
	}

	memset(instance->effects, 0, instance->n_effects * sizeof *instance->effects);
	return 0;
}

int bfm_instance_destroy(bfm_instance_t* instance) {
	bfm_state_t* const state = instance->state;

	if (instance->effects != NULL) {
		bfmg_host_unpin(instance->effects); /* bfm_sim_run page-locks large result buffers in place */
		state->free(instance->effects);
	}

	if (instance->conditions != NULL) {
		state->free(instance->conditions);
	}

	return 0;
}

/* shared by the instance/sim pointer lists: replace the list by n zeroed slots */
static int ptr_list_reset(bfm_state_t* state, void*** list, size_t* count, size_t n) {
	*count = n;

	if (*list != NULL) {
		state->free(*list);
	}

	*list = state->alloc(n * sizeof **list);

	if (*list == NULL) {
		return -1;
	}

	memset(*list, 0, n * sizeof **list);
	return 0;
}

/* ... and append one borrowed pointer */
static int ptr_list_push(bfm_state_t* state, void*** list, size_t* count, void* item) {
	*list = state->realloc(*list, ++*count * sizeof **list);

	if (*list == NULL) {
		return -1;
	}

	(*list)[*count - 1] = item;
	return 0;
}

int bfm_instance_set_n_conditions(bfm_instance_t* instance, size_t n_conditions) {
	return ptr_list_reset(instance->state, (void***) &instance->conditions, &instance->n_conditions, n_conditions);
}

int bfm_instance_add_condition(bfm_instance_t* instance, bfm_condition_t* condition) {
	return ptr_list_push(instance->state, (void***) &instance->conditions, &instance->n_conditions, condition);
}

/* ---- forces (reference force.c:5-99) --------------------------------------------------------- */

int bfm_force_create(bfm_force_t* force, bfm_state_t* state, size_t dim) {
	memset(force, 0, sizeof *force);

	force->state = state;
	force->dim = dim;

	return 0;
}

int bfm_force_destroy(bfm_force_t* force) {
	(void) force; /* the deep copy made by set_linear is never released (force.c:14-18) - kept,
	               * because callers such as bfm_ez_lepl1110_destroy may destroy a zeroed force */
	return 0;
}

int bfm_force_set_none(bfm_force_t* force) {
	force->kind = BFM_FORCE_KIND_NONE;
	return 0;
}

int bfm_force_set_linear(bfm_force_t* force, bfm_vec_t* vec) {
	force->kind = BFM_FORCE_KIND_LINEAR;

	if (vec->n != force->dim) {
		return -1;
	}

	if (bfm_vec_create(&force->linear.force, vec->state, vec->n) < 0) {
		return -1;
	}

	return bfm_vec_copy(&force->linear.force, vec);
}

int bfm_force_set_funky(bfm_force_t* force, bfm_force_funky_func_t func, void* data) {
	force->kind = BFM_FORCE_KIND_FUNKY;
	force->funky.func = func;
	force->funky.data = data;

	return 0;
}

int bfm_force_eval(bfm_force_t* force, bfm_vec_t* pos, bfm_vec_t* force_ref) {
	if (force_ref->n != force->dim) {
		return -1;
	}

	switch (force->kind) {
	case BFM_FORCE_KIND_NONE:
		memset(force_ref->data, 0, force_ref->n * sizeof *force_ref->data);
		return 0;

	case BFM_FORCE_KIND_LINEAR:
		bfm_vec_copy(force_ref, &force->linear.force);
		return -1; /* sic: the reference's linear evaluator reports -1 after a successful copy
		            * (force.c:65-73) and its only caller ignores the status (system.c:198) */

	case BFM_FORCE_KIND_FUNKY:
		return force->funky.func(force, pos, force_ref, force->funky.data);
	}

	return -1;
}

/* ---- simulation container (reference sim.c:8-97); bfm_sim_run lives in sim.c ----------------- */

int bfm_sim_create(bfm_sim_t* sim, bfm_state_t* state, bfm_sim_kind_t kind) {
	memset(sim, 0, sizeof *sim);

	sim->state = state;
	sim->kind = kind;

	return 0;
}

int bfm_sim_destroy(bfm_sim_t* sim) {
	bfm_state_t* const state = sim->state;

	if (sim->instances != NULL) {
		state->free(sim->instances);
	}

	if (sim->forces != NULL) {
		state->free(sim->forces);
	}

	return 0;
}

int bfm_sim_set_n_instances(bfm_sim_t* sim, size_t n_instances) {
	return ptr_list_reset(sim->state, (void***) &sim->instances, &sim->n_instances, n_instances);
}

int bfm_sim_add_instance(bfm_sim_t* sim, bfm_instance_t* instance) {
	return ptr_list_push(sim->state, (void***) &sim->instances, &sim->n_instances, instance);
}

int bfm_sim_set_n_forces(bfm_sim_t* sim, size_t n_forces) {
	return ptr_list_reset(sim->state, (void***) &sim->forces, &sim->n_forces, n_forces);
}

int bfm_sim_add_force(bfm_sim_t* sim, bfm_force_t* force) {
	return ptr_list_push(sim->state, (void***) &sim->forces, &sim->n_forces, force);
}
