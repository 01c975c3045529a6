/*
 * ez.c - LEPL1110 convenience layer: builds a complete simulation (material, rule, object,
 * instance, gravity, boundary conditions) from a course "problem file" and writes the U/V result
 * files.  Replaces reference ez.c:6-229; config 1's entry point.
 *
 * Problem file grammar (problems/problem.txt), one "key : value" per line, key padded to 19 columns
 * and compared case-insensitively on those 19 columns:
 *
 *   Type of problem    :  Planar strains | Planar stresses | Axi-symetric problem
 *   Young modulus      :  <E>            Poisson ratio      :  <nu>
 *   Mass density       :  <rho>          Gravity            :  <g>       (force (0, -g))
 *   Boundary condition :  <kind> = <value> : <domain name>
 */
#include "internal.h"

#include <stdio.h>
#include <string.h>
#include <strings.h>

static struct {
	char const* name;
	bfm_condition_kind_t kind;
} const condition_names[] = {
	{"Dirichlet-X", BFM_CONDITION_KIND_DIRICHLET_X},
	{"Dirichlet-Y", BFM_CONDITION_KIND_DIRICHLET_Y},
	{"Neumann-X", BFM_CONDITION_KIND_NEUMANN_X},
	{"Neumann-Y", BFM_CONDITION_KIND_NEUMANN_Y},
	{"Neumann-Tangent", BFM_CONDITION_KIND_NEUMANN_TANGENT},
	{"Neumann-Normal", BFM_CONDITION_KIND_NEUMANN_NORMAL},
	{"Dirichlet-Normal", BFM_CONDITION_KIND_DIRICHLET_NORMAL},
	{"Dirichlet-Tangent", BFM_CONDITION_KIND_DIRICHLET_TANGENT},
};

static bool key_is(char const* line, char const* key) {
	return strncasecmp(line, key, 19) == 0;
}

/* text after the first ':' of the line, leading blanks removed, newline cut */
static char* value_of(char* line) {
	char* v = strchr(line, ':');

	if (v == NULL) {
		return NULL;
	}

	for (v++; *v == ' ' || *v == '\t'; v++) {
	}

	v[strcspn(v, "\r\n")] = '\0';
	return v;
}

static int add_gravity(bfm_ez_lepl1110_t* ez, double g) {
	bfm_state_t* const state = ez->state;
	bfm_vec_t down;

	if (bfm_force_create(&ez->gravity, state, 2) < 0 || bfm_vec_create(&down, state, 2) < 0) {
		return -1;
	}

	down.data[1] = g;
	down.data[1] *= -1; /* ez.c:97-98: read g, then flip it */

	int const rv = bfm_force_set_linear(&ez->gravity, &down);
	bfm_vec_destroy(&down);

	if (rv < 0) {
		return -1;
	}

	return bfm_sim_add_force(&ez->sim, &ez->gravity);
}

static int add_condition(bfm_ez_lepl1110_t* ez, char* spec) {
	bfm_state_t* const state = ez->state;
	bfm_mesh_t* const mesh = ez->mesh;

	/* "<kind> = <value> : <domain name>" */

	char kind_str[20];
	double value;
	int used = 0;

	if (sscanf(spec, "%19s = %le : %n", kind_str, &value, &used) < 2 || used == 0) {
		return -1;
	}

	char const* const domain_name = spec + used;
	bfm_condition_kind_t kind = 0;
	bool known = false;

	for (size_t i = 0; i < sizeof condition_names / sizeof *condition_names; i++) {
		if (strncasecmp(kind_str, condition_names[i].name, 19) == 0) {
			kind = condition_names[i].kind;
			known = true;
			break;
		}
	}

	if (!known) {
		return -1; /* ez.c:145-147 */
	}

	ez->conditions = state->realloc(ez->conditions, ++ez->n_conditions * sizeof *ez->conditions);

	if (ez->conditions == NULL) {
		return -1;
	}

	bfm_condition_t* const condition = &ez->conditions[ez->n_conditions - 1];

	if (bfm_condition_create(condition, state, mesh, kind) < 0) {
		return -1;
	}

	condition->value = value;

	/* the first domain whose name matches on 25 characters marks both end nodes of each of its
	 * edges (ez.c:158-173) */

	for (size_t i = 0; i < mesh->n_domains; i++) {
		bfm_domain_t const* const domain = &mesh->domains[i];

		if (strncasecmp(domain->name, domain_name, 25) != 0) {
			continue;
		}

		for (size_t j = 0; j < domain->n_elements; j++) {
			bfm_edge_t const* const edge = &mesh->edges[domain->elements[j]];

			condition->nodes[edge->nodes[0]] = true;
			condition->nodes[edge->nodes[1]] = true;
		}

		break;
	}

	return 0;
}

int bfm_ez_lepl1110_create(bfm_ez_lepl1110_t* ez, bfm_state_t* state, bfm_mesh_t* mesh, char* name) {
	ez->state = state;
	ez->mesh = mesh;

	/* placeholders first, the problem file overwrites them (ez.c:13-36) */

	if (
		bfm_sim_create(&ez->sim, state, BFM_SIM_KIND_NONE) < 0 ||
		bfm_material_create(&ez->material, state, "lepl1110", 0, 0, 0) < 0 ||
		bfm_rule_create_gauss_legendre(&ez->rule, state, 2, mesh->kind) < 0 ||
		bfm_obj_create(&ez->obj, state, mesh, &ez->material, &ez->rule) < 0 ||
		bfm_instance_create(&ez->instance, state, &ez->obj) < 0
	) {
		return -1;
	}

	bfm_sim_add_instance(&ez->sim, &ez->instance);

	FILE* const fp = fopen(name, "r");

	if (fp == NULL) {
		return -1;
	}

	char line[256];
	int rv = 0;

	while (rv == 0 && fgets(line, sizeof line, fp) != NULL) {
		char* const value = value_of(line);

		if (value == NULL) {
			continue;
		}

		if (key_is(line, "Type of problem     ")) {
			if (strncasecmp(value, "Planar strains", 13) == 0) {
				ez->sim.kind = BFM_SIM_KIND_PLANAR_STRAIN;
			}

			else if (strncasecmp(value, "Planar stresses", 13) == 0) {
				ez->sim.kind = BFM_SIM_KIND_PLANAR_STRESS;
			}

			else if (strncasecmp(value, "Axi-symetric problem", 13) == 0) {
				ez->sim.kind = BFM_SIM_KIND_AXISYMMETRIC_STRAIN;
			}
		}

		else if (key_is(line, "Young modulus       ")) {
			sscanf(value, "%le", &ez->material.E);
		}

		else if (key_is(line, "Poisson ratio       ")) {
			sscanf(value, "%le", &ez->material.nu);
		}

		else if (key_is(line, "Mass density        ")) {
			sscanf(value, "%le", &ez->material.rho);
		}

		else if (key_is(line, "Gravity             ")) {
			double g = 0;
			sscanf(value, "%le", &g);
			rv = add_gravity(ez, g);
		}

		else if (key_is(line, "Boundary condition  ")) {
			rv = add_condition(ez, value);
		}
	}

	fclose(fp);

	if (rv < 0) {
		return -1;
	}

	/* register the conditions only now: the array above moves while it grows (ez.c:181-186) */

	for (size_t i = 0; i < ez->n_conditions; i++) {
		if (bfm_instance_add_condition(&ez->instance, &ez->conditions[i]) < 0) {
			return -1;
		}
	}

	return 0;
}

int bfm_ez_lepl1110_destroy(bfm_ez_lepl1110_t* ez) {
	bfm_state_t* const state = ez->state;

	for (size_t i = 0; i < ez->n_conditions; i++) {
		bfm_condition_destroy(&ez->conditions[i]);
	}

	state->free(ez->conditions);

	bfm_force_destroy(&ez->gravity);
	bfm_material_destroy(&ez->material);
	bfm_rule_destroy(&ez->rule);
	bfm_obj_destroy(&ez->obj);
	bfm_instance_destroy(&ez->instance);
	bfm_sim_destroy(&ez->sim);

	return 0;
}

/* reference ez.c:211-229: "%14.7e", a line break after every third value */
int bfm_ez_lepl1110_write(bfm_ez_lepl1110_t* ez, size_t shift, char const* filename) {
	FILE* const fp = fopen(filename, "w");

	if (fp == NULL) {
		return -1;
	}

	size_t const n_nodes = ez->obj.mesh->n_nodes;

	fprintf(fp, "Number of nodes %zu\n", n_nodes);

	for (size_t i = 0; i < n_nodes; i++) {
		fprintf(fp, "%14.7e", ez->instance.effects[i * 2 + shift]);

		if ((i + 1) % 3 == 0 && i + 1 != ez->instance.n_effects) {
			fputc('\n', fp);
		}
	}

	fputc('\n', fp);
	fclose(fp);

	return 0;
}
