/*
 * partition.c - row partition of a mesh over the ranks of one multi-GPU job (host side, no device code).
 *
 * The reference is a single-threaded dense solver with nothing to partition; this
 * is the new capability BASELINE.json's north star asks for: "large meshes are row-partitioned across
 * the GPUs of one 8xB200 box".  Rank r owns the contiguous global node range
 * [r * n / world, (r + 1) * n / world) - hence both DOFs of a node, and a contiguous block of matrix rows.
 *
 * A rank's LOCAL mesh holds every element that touches one of its nodes, so each owned matrix row sees
 * all of its contributions and needs no assembly communication (elements on a cut are computed by
 * both sides).  Local node ids are the ascending global ids of (owned + ghost) nodes and local element
 * ids the ascending global ids of the selected elements: the numbering is MONOTONE, so every order the
 * reference's arithmetic depends on - elements ascending inside an entry (system.c:460), columns
 * ascending inside a row (system.c:358-374), DOFs ascending inside a condition (system.c:376-386) -
 * is the same on the local mesh as on the global one, and the owned rows come out bit-identical to a
 * single-GPU assembly.  Owned nodes form the contiguous local range [own_begin, own_end); ghost rows
 * are assembled only partially and never used.
 *
 * Halo plan: the ghosts owned by one neighbour are a contiguous local range (owners are contiguous in
 * global ids), and the list of owned nodes a neighbour needs follows from the local elements alone -
 * node a goes to rank s iff it shares an element with a node of s, a relation both sides see - so the
 * two sides agree on contents and order (ascending global id) without exchanging anything.
 */
#include "internal.h"

#include <stdlib.h>
#include <string.h>

static int cmp_size(void const* a, void const* b) {
	size_t const x = *(size_t const*) a;
	size_t const y = *(size_t const*) b;
	return x < y ? -1 : x > y;
}

static int cmp_u64(void const* a, void const* b) {
	uint64_t const x = *(uint64_t const*) a;
	uint64_t const y = *(uint64_t const*) b;
	return x < y ? -1 : x > y;
}

size_t bfmi_part_first_node(size_t n_nodes, int world, int rank) {
	return (size_t) (((unsigned __int128) n_nodes * (unsigned) rank) / (unsigned) world);
}

int bfmi_part_owner(size_t n_nodes, int world, size_t node) {
	int r = (int) (((unsigned __int128) node * (unsigned) world) / n_nodes);

	r = r >= world ? world - 1 : r;

	while (r > 0 && node < bfmi_part_first_node(n_nodes, world, r)) {
		r--;
	}

	while (r + 1 < world && node >= bfmi_part_first_node(n_nodes, world, r + 1)) {
		r++;
	}

	return r;
}

int32_t bfmi_part_local(bfmi_part_t const* part, size_t g) {
	if (g >= part->lo && g < part->hi) {
		return part->own_begin + (int32_t) (g - part->lo);
	}

	/* ghosts: binary search in the ascending local-to-global table on the matching side */

	int32_t lo = g < part->lo ? 0 : part->own_end;
	int32_t hi = g < part->lo ? part->own_begin : part->n_local;

	while (lo < hi) {
		int32_t const mid = lo + (hi - lo) / 2;

		if (part->l2g[mid] < g) {
			lo = mid + 1;
		}

		else {
			hi = mid;
		}
	}

	return lo < part->n_local && part->l2g[lo] == g ? lo : -1;
}

static void part_free(bfmi_part_t* part) {
	if (part == NULL) {
		return;
	}

	bfmi_plan_forget(&part->local);
	bfmg_free(part->d_send_idx);

	bfmg_host_unpin(part->local.coords);

	free(part->l2g);
	free(part->elem_l2g);
	free(part->local.coords);
	free(part->local.elems);
	free(part->nbr);
	free(part->recv_begin);
	free(part->recv_count);
	free(part->send_ptr);
	free(part->send_idx);
	free(part);
}

void bfmi_part_release(bfmi_part_t* part) {
	if (part != NULL && __atomic_sub_fetch(&part->refs, 1, __ATOMIC_ACQ_REL) == 0) {
		part_free(part);
	}
}

bfmi_part_t* bfmi_part_build(bfm_state_t* state, bfm_mesh_t const* mesh, int rank, int world) {
	size_t const nn = mesh->n_nodes;
	size_t const ne = mesh->n_elems;
	size_t const kind = mesh->kind;

	if (world < 1 || rank < 0 || rank >= world || nn < (size_t) world) {
		BFMI_FAIL(state, "cannot partition %zu nodes over %d ranks", nn, world);
		return NULL;
	}

	bfmi_part_t* const part = calloc(1, sizeof *part);

	if (part == NULL) {
		return NULL;
	}

	part->refs = 1;
	part->global = mesh;
	part->n_nodes = nn;
	part->n_elems = ne;
	part->rank = rank;
	part->world = world;
	part->lo = bfmi_part_first_node(nn, world, rank);
	part->hi = bfmi_part_first_node(nn, world, rank + 1);

	size_t const lo = part->lo;
	size_t const hi = part->hi;

	/* pass 1: select elements, collect ghost candidates */

	size_t n_sel = 0;
	size_t n_cand = 0;

	for (size_t e = 0; e < ne; e++) {
		size_t owned = 0;

		for (size_t j = 0; j < kind; j++) {
			size_t const g = mesh->elems[e * kind + j];

			if (g >= nn) {
				BFMI_FAIL(state, "element %zu points outside the node table", e);
				goto fail;
			}

			owned += g >= lo && g < hi;
		}

		if (owned) {
			n_sel++;
			n_cand += kind - owned;
		}
	}

	part->elem_l2g = malloc((n_sel + 1) * sizeof *part->elem_l2g);
	size_t* cand = malloc((n_cand + 1) * sizeof *cand);

	if (part->elem_l2g == NULL || cand == NULL) {
		free(cand);
		goto fail;
	}

	n_sel = n_cand = 0;

	for (size_t e = 0; e < ne; e++) {
		bool any = false;

		for (size_t j = 0; j < kind; j++) {
			size_t const g = mesh->elems[e * kind + j];
			any |= g >= lo && g < hi;
		}

		if (!any) {
			continue;
		}

		part->elem_l2g[n_sel++] = e;

		for (size_t j = 0; j < kind; j++) {
			size_t const g = mesh->elems[e * kind + j];

			if (g < lo || g >= hi) {
				cand[n_cand++] = g;
			}
		}
	}

	qsort(cand, n_cand, sizeof *cand, cmp_size);

	size_t n_ghost = 0;
	size_t n_below = 0;

	for (size_t i = 0; i < n_cand; i++) {
		if (i == 0 || cand[i] != cand[i - 1]) {
			cand[n_ghost++] = cand[i];
			n_below += cand[i] < lo;
		}
	}

	size_t const n_local = (hi - lo) + n_ghost;

	if (n_local >= (1u << 30)) {
		free(cand);
		goto fail;
	}

	part->n_local = (int32_t) n_local;
	part->own_begin = (int32_t) n_below;
	part->own_end = (int32_t) (n_below + (hi - lo));
	part->l2g = malloc((n_local + 1) * sizeof *part->l2g);

	if (part->l2g == NULL) {
		free(cand);
		goto fail;
	}

	memcpy(part->l2g, cand, n_below * sizeof *cand);

	for (size_t g = lo; g < hi; g++) {
		part->l2g[n_below + (g - lo)] = g;
	}

	memcpy(part->l2g + part->own_end, cand + n_below, (n_ghost - n_below) * sizeof *cand);
	free(cand);

	/* the local mesh: coordinates and connectivity in local ids (plain malloc: it never meets
	 * bfm_mesh_destroy); edges stay global - the boundary-condition lists are built on the global mesh
	 * and renumbered (job.c) */

	bfm_mesh_t* const local = &part->local;

	local->state = mesh->state;
	local->dim = mesh->dim;
	local->kind = mesh->kind;
	local->n_nodes = n_local;
	local->n_elems = n_sel;
	local->coords = malloc((n_local * 2 + 1) * sizeof *local->coords);
	local->elems = malloc((n_sel * kind + 1) * sizeof *local->elems);

	if (local->coords == NULL || local->elems == NULL) {
		goto fail;
	}

#pragma omp parallel for schedule(static) if (n_local > 100000)
	for (size_t l = 0; l < n_local; l++) {
		local->coords[2 * l + 0] = mesh->coords[2 * part->l2g[l] + 0];
		local->coords[2 * l + 1] = mesh->coords[2 * part->l2g[l] + 1];
	}

#pragma omp parallel for schedule(static) if (n_sel > 100000)
	for (size_t le = 0; le < n_sel; le++) {
		for (size_t j = 0; j < kind; j++) {
			local->elems[le * kind + j] = (size_t) bfmi_part_local(part, mesh->elems[part->elem_l2g[le] * kind + j]);
		}
	}

	/* halo plan.  Receive side: ghost ranges by owner.  Send side: (neighbour, owned local node) pairs
	 * gathered from the elements that straddle a cut. */

	int* const is_nbr = calloc((size_t) world, sizeof *is_nbr);
	uint64_t* pairs = NULL;
	size_t n_pairs = 0;
	size_t cap_pairs = 0;

	if (is_nbr == NULL) {
		goto fail;
	}

	for (int32_t l = 0; l < part->n_local; l++) {
		if (l < part->own_begin || l >= part->own_end) {
			is_nbr[bfmi_part_owner(nn, world, part->l2g[l])] = 1;
		}
	}

	for (int r = 0; r < world; r++) {
		part->n_nbr += is_nbr[r];
	}

	free(is_nbr);

	part->nbr = malloc(((size_t) part->n_nbr + 1) * sizeof *part->nbr);
	part->recv_begin = malloc(((size_t) part->n_nbr + 1) * sizeof *part->recv_begin);
	part->recv_count = calloc((size_t) part->n_nbr + 1, sizeof *part->recv_count);
	part->send_ptr = calloc((size_t) part->n_nbr + 2, sizeof *part->send_ptr);

	if (part->nbr == NULL || part->recv_begin == NULL || part->recv_count == NULL || part->send_ptr == NULL) {
		goto fail;
	}

	{
		int idx = -1;
		int prev = -1;

		for (int32_t l = 0; l < part->n_local; l++) {
			if (l >= part->own_begin && l < part->own_end) {
				continue;
			}

			int const owner = bfmi_part_owner(nn, world, part->l2g[l]);

			if (owner != prev) { /* ghosts ascend in global id, so owners ascend too */
				idx++;
				part->nbr[idx] = owner;
				part->recv_begin[idx] = l;
				prev = owner;
			}

			part->recv_count[idx]++;
		}
	}

	for (size_t le = 0; le < n_sel; le++) {
		size_t const* const el = &local->elems[le * kind];
		bool cut = false;

		for (size_t j = 0; j < kind; j++) {
			cut |= (int32_t) el[j] < part->own_begin || (int32_t) el[j] >= part->own_end;
		}

		if (!cut) {
			continue;
		}

		if (n_pairs + kind * kind > cap_pairs) {
			cap_pairs = cap_pairs ? 2 * cap_pairs : 4096;
			uint64_t* const grown = realloc(pairs, cap_pairs * sizeof *pairs);

			if (grown == NULL) {
				free(pairs);
				goto fail;
			}

			pairs = grown;
		}

		for (size_t j = 0; j < kind; j++) {
			if ((int32_t) el[j] < part->own_begin || (int32_t) el[j] >= part->own_end) {
				continue; /* not mine to send */
			}

			for (size_t k = 0; k < kind; k++) {
				if ((int32_t) el[k] >= part->own_begin && (int32_t) el[k] < part->own_end) {
					continue;
				}

				uint64_t const owner = (uint64_t) bfmi_part_owner(nn, world, part->l2g[el[k]]);
				pairs[n_pairs++] = owner << 32 | (uint64_t) el[j];
			}
		}
	}

	if (n_pairs > 0) {
		qsort(pairs, n_pairs, sizeof *pairs, cmp_u64);
	}

	size_t n_send = 0;

	for (size_t i = 0; i < n_pairs; i++) {
		if (i == 0 || pairs[i] != pairs[i - 1]) {
			pairs[n_send++] = pairs[i];
		}
	}

	part->send_idx = malloc((n_send + 1) * sizeof *part->send_idx);

	if (part->send_idx == NULL) {
		free(pairs);
		goto fail;
	}

	for (size_t i = 0; i < n_send; i++) {
		int const owner = (int) (pairs[i] >> 32);
		int idx = 0;

		while (idx < part->n_nbr && part->nbr[idx] != owner) {
			idx++;
		}

		if (idx == part->n_nbr) { /* cannot happen: a node I send to owns a node I ghost */
			free(pairs);
			goto fail;
		}

		part->send_ptr[idx + 1]++;
		part->send_idx[i] = (int32_t) (pairs[i] & 0xffffffffu);
	}

	free(pairs);

	for (int i = 0; i < part->n_nbr; i++) {
		part->send_ptr[i + 1] += part->send_ptr[i];
	}

	part->n_send = (int32_t) n_send;
	return part;

fail:

	part_free(part);
	return NULL;
}

/* ---- cache: one partition per (mesh, rank, world), keyed like the plan cache ---------------------- */

static bfmi_part_t* cached;

bfmi_part_t* bfmi_part_for_mesh(bfm_state_t* state, bfm_mesh_t const* mesh, uint64_t hash, int rank, int world) {
	if (cached != NULL && cached->global == mesh && cached->n_nodes == mesh->n_nodes && cached->n_elems == mesh->n_elems && cached->hash == hash && cached->rank == rank && cached->world == world) {
		/* same connectivity: the partition stands, but the caller may have moved the nodes since */

#pragma omp parallel for schedule(static) if (cached->n_local > 100000)
		for (int32_t l = 0; l < cached->n_local; l++) {
			cached->local.coords[2 * l + 0] = mesh->coords[2 * cached->l2g[l] + 0];
			cached->local.coords[2 * l + 1] = mesh->coords[2 * cached->l2g[l] + 1];
		}

		__atomic_add_fetch(&cached->refs, 1, __ATOMIC_RELAXED);
		return cached;
	}

	bfmi_part_t* const part = bfmi_part_build(state, mesh, rank, world);

	if (part == NULL) {
		return NULL;
	}

	part->hash = hash;

	bfmi_part_release(cached);
	cached = part;

	__atomic_add_fetch(&part->refs, 1, __ATOMIC_RELAXED);
	return part;
}

void bfmi_part_forget(bfm_mesh_t const* mesh) {
	if (cached != NULL && cached->global == mesh) {
		bfmi_part_release(cached);
		cached = NULL;
	}
}

/* ---- introspection (bfm_b200.h): what the gloo tests and a curious caller look at ------------------- */

int bfmx_partition_sizes(bfm_mesh_t* mesh, int rank, int world, bfmx_partition_info_t* out) {
	bfmi_part_t* const part = bfmi_part_build(mesh->state, mesh, rank, world);

	if (part == NULL) {
		return -1;
	}

	out->first_node = part->lo;
	out->end_node = part->hi;
	out->n_local_nodes = (size_t) part->n_local;
	out->own_begin = (size_t) part->own_begin;
	out->own_end = (size_t) part->own_end;
	out->n_local_elems = part->local.n_elems;
	out->n_neighbours = (size_t) part->n_nbr;
	out->n_send = (size_t) part->n_send;

	bfmi_part_release(part);
	return 0;
}

int bfmx_partition_copy(bfm_mesh_t* mesh, int rank, int world, size_t* local_to_global, size_t* local_elems, size_t* elem_to_global, int32_t* neighbours, int32_t* recv_begin, int32_t* recv_count, int32_t* send_ptr, int32_t* send_idx) {
	bfmi_part_t* const part = bfmi_part_build(mesh->state, mesh, rank, world);

	if (part == NULL) {
		return -1;
	}

	memcpy(local_to_global, part->l2g, (size_t) part->n_local * sizeof *local_to_global);
	memcpy(local_elems, part->local.elems, part->local.n_elems * part->local.kind * sizeof *local_elems);
	memcpy(elem_to_global, part->elem_l2g, part->local.n_elems * sizeof *elem_to_global);
	memcpy(neighbours, part->nbr, (size_t) part->n_nbr * sizeof *neighbours);
	memcpy(recv_begin, part->recv_begin, (size_t) part->n_nbr * sizeof *recv_begin);
	memcpy(recv_count, part->recv_count, (size_t) part->n_nbr * sizeof *recv_count);
	memcpy(send_ptr, part->send_ptr, ((size_t) part->n_nbr + 1) * sizeof *send_ptr);
	memcpy(send_idx, part->send_idx, (size_t) part->n_send * sizeof *send_idx);

	bfmi_part_release(part);
	return 0;
}
