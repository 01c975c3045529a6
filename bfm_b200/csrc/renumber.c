/*
 * renumber.c - internal node numbering for meshes whose own numbering has no locality.
 *
 * The reference stores the system densely and then looks for a good numbering itself (bfm_system_renumber: RCM,
 * system.c:44-81, perm.c:117-331) because its band LU needs one.  The sparse kernels here need something weaker but
 * need it badly: while a warp streams 32 consecutive block rows, the entries of the vector it gathers must sit in
 * L2.  With the numbering a mesh generator produces they do; with an arbitrary one they do not - measured on the
 * B200 with a randomly numbered 18 M-DOF plate: CG's SpMV at 2.8 instead of 6.7 TB/s, the solve 2.1x slower
 * (profiles/r2_summary.md) - and a row partition over several GPUs of such a mesh is all halo.
 *
 * So bfm_sim_run looks at the numbering first: the mean over the elements of (largest - smallest node number) is
 * the half-width of the window of the gathered vector that has to stay cached; when it exceeds 2^20 nodes
 * (32 MB of vector: a quarter of the 126 MB L2) the job runs on an internal copy of the mesh whose nodes are numbered
 * along a Morton curve through their coordinates.  What does not change: the elements and their order, and the
 * order of the nodes inside an element - so every matrix block and every load-vector entry is accumulated from the
 * same contributions in the same order and keeps its bits; which DOFs a condition constrains.  What may change in
 * the last bit: sums whose order follows the numbering (the right-hand-side corrections of a row next to several
 * non-zero Dirichlet values, the dot products of CG).  Displacements are written back in the caller's numbering.
 * The public matrix API (bfm_system_create_*: values and pattern are compared bit for bit) never renumbers.
 *
 * BFM_RENUMBER=0 keeps the caller's numbering whatever it is, BFM_RENUMBER=1 renumbers every mesh bfm_sim_run sees.
 * Cached per mesh object + connectivity hash (one slot), like the partition; the coordinates are refreshed on every
 * call (the caller may have moved the nodes).
 */
#include "internal.h"

#include <stdlib.h>
#include <string.h>

static void renum_free(bfmi_renum_t* r) {
	if (r == NULL) {
		return;
	}

	bfmi_plan_forget(&r->mesh);
	bfmi_part_forget(&r->mesh);
	bfmi_coarse_forget(&r->mesh);

	bfmg_host_unpin(r->mesh.coords);
	bfmg_free(r->d_to_new);

	free(r->mesh.coords);
	free(r->mesh.elems);
	free(r->to_new);
	free(r->to_old);
	free(r);
}

void bfmi_renum_release(bfmi_renum_t* r) {
	if (r != NULL && __atomic_sub_fetch(&r->refs, 1, __ATOMIC_ACQ_REL) == 0) {
		renum_free(r);
	}
}

/* mean over the elements of (largest - smallest node number) */
static double mean_span(bfm_mesh_t const* mesh) {
	size_t const kind = mesh->kind;
	size_t const ne = mesh->n_elems;
	double total = 0;

#pragma omp parallel for schedule(static) reduction(+ : total) if (ne > 100000)
	for (size_t e = 0; e < ne; e++) {
		size_t lo = mesh->elems[e * kind], hi = lo;

		for (size_t j = 1; j < kind; j++) {
			size_t const v = mesh->elems[e * kind + j];
			lo = v < lo ? v : lo;
			hi = v > hi ? v : hi;
		}

		total += (double) (hi - lo);
	}

	return ne > 0 ? total / (double) ne : 0;
}

/* bits of x and y interleaved (x in the even positions) */
static inline uint32_t spread16(uint32_t v) {
	v &= 0xffffu;
	v = (v | v << 8) & 0x00ff00ffu;
	v = (v | v << 4) & 0x0f0f0f0fu;
	v = (v | v << 2) & 0x33333333u;
	v = (v | v << 1) & 0x55555555u;
	return v;
}

/* to_old[new] = the caller's node at position `new` of the Morton order (ties: ascending caller number) */
static int morton_order(bfm_mesh_t const* mesh, int32_t* to_old) {
	size_t const nn = mesh->n_nodes;
	double lo[2] = {mesh->coords[0], mesh->coords[1]};
	double hi[2] = {mesh->coords[0], mesh->coords[1]};

	for (size_t a = 0; a < nn; a++) {
		for (int c = 0; c < 2; c++) {
			double const v = mesh->coords[2 * a + c];
			lo[c] = v < lo[c] ? v : lo[c];
			hi[c] = v > hi[c] ? v : hi[c];
		}
	}

	/* square cells: both axes are quantised with the step of the longer one */
	double const extent = hi[0] - lo[0] > hi[1] - lo[1] ? hi[0] - lo[0] : hi[1] - lo[1];
	double const scale = extent > 0 ? 65535.0 / extent : 0;

	uint32_t* const key = malloc((nn + 1) * sizeof *key);
	int32_t* const tmp = malloc((nn + 1) * sizeof *tmp);
	size_t* const count = malloc(65537 * sizeof *count);

	if (key == NULL || tmp == NULL || count == NULL) {
		free(key), free(tmp), free(count);
		return -1;
	}

#pragma omp parallel for schedule(static) if (nn > 100000)
	for (size_t a = 0; a < nn; a++) {
		double const qx = (mesh->coords[2 * a + 0] - lo[0]) * scale;
		double const qy = (mesh->coords[2 * a + 1] - lo[1]) * scale;
		uint32_t const ix = qx >= 0 && qx < 65536 ? (uint32_t) qx : 0; /* NaN and friends land in cell 0 */
		uint32_t const iy = qy >= 0 && qy < 65536 ? (uint32_t) qy : 0;

		key[a] = spread16(ix) | spread16(iy) << 1;
	}

	/* stable radix sort of the node numbers by key, 16 bits per pass */

	for (int pass = 0; pass < 2; pass++) {
		int const shift = 16 * pass;
		int32_t const* const src = pass == 0 ? NULL : tmp;
		int32_t* const dst = pass == 0 ? tmp : to_old;

		memset(count, 0, 65537 * sizeof *count);

		for (size_t i = 0; i < nn; i++) {
			size_t const a = src != NULL ? (size_t) src[i] : i;
			count[(key[a] >> shift & 0xffffu) + 1]++;
		}

		for (size_t b = 0; b < 65536; b++) {
			count[b + 1] += count[b];
		}

		for (size_t i = 0; i < nn; i++) {
			size_t const a = src != NULL ? (size_t) src[i] : i;
			dst[count[key[a] >> shift & 0xffffu]++] = (int32_t) a;
		}
	}

	free(key), free(tmp), free(count);
	return 0;
}

static bfmi_renum_t* renum_build(bfm_mesh_t const* mesh, uint64_t hash) {
	size_t const nn = mesh->n_nodes;
	size_t const n_conn = mesh->n_elems * mesh->kind;

	if (nn == 0 || nn >= (1u << 30)) {
		return NULL;
	}

	for (size_t i = 0; i < n_conn; i++) {
		if (mesh->elems[i] >= nn) {
			return NULL; /* the plan builder reports it */
		}
	}

	bfmi_renum_t* const r = calloc(1, sizeof *r);

	if (r == NULL) {
		return NULL;
	}

	r->refs = 1;
	r->orig = mesh;
	r->n_nodes = nn;
	r->n_elems = mesh->n_elems;
	r->kind = (int) mesh->kind;
	r->hash = hash;

	r->to_new = malloc((nn + 1) * sizeof *r->to_new);
	r->to_old = malloc((nn + 1) * sizeof *r->to_old);

	r->mesh.state = mesh->state;
	r->mesh.dim = mesh->dim;
	r->mesh.kind = mesh->kind;
	r->mesh.n_nodes = nn;
	r->mesh.n_elems = mesh->n_elems;
	r->mesh.coords = malloc((nn * 2 + 1) * sizeof *r->mesh.coords);
	r->mesh.elems = malloc((n_conn + 1) * sizeof *r->mesh.elems);

	if (r->to_new == NULL || r->to_old == NULL || r->mesh.coords == NULL || r->mesh.elems == NULL || morton_order(mesh, r->to_old) < 0) {
		renum_free(r);
		return NULL;
	}

#pragma omp parallel for schedule(static) if (nn > 100000)
	for (size_t i = 0; i < nn; i++) {
		r->to_new[r->to_old[i]] = (int32_t) i;
	}

	/* the same elements in the same order, the same node order inside each */

#pragma omp parallel for schedule(static) if (n_conn > 100000)
	for (size_t i = 0; i < n_conn; i++) {
		r->mesh.elems[i] = (size_t) r->to_new[mesh->elems[i]];
	}

	r->mesh_hash = bfmi_mesh_hash(&r->mesh);
	return r;
}

static void refresh_coords(bfmi_renum_t* r, bfm_mesh_t const* mesh) {
#pragma omp parallel for schedule(static) if (r->n_nodes > 100000)
	for (size_t i = 0; i < r->n_nodes; i++) {
		r->mesh.coords[2 * i + 0] = mesh->coords[2 * (size_t) r->to_old[i] + 0];
		r->mesh.coords[2 * i + 1] = mesh->coords[2 * (size_t) r->to_old[i] + 1];
	}
}

/* one slot: the renumbered copy of the last mesh that needed one; and the last mesh found not to need one */
static bfmi_renum_t* cached;

static struct {
	bfm_mesh_t const* mesh;
	size_t n_nodes, n_elems;
	uint64_t hash;
} kept;

bfmi_renum_t* bfmi_renum_for_mesh(bfm_mesh_t const* mesh, uint64_t hash) {
	char const* const env = getenv("BFM_RENUMBER");
	int const mode = env != NULL && env[0] != 0 ? atoi(env) : -1; /* -1: decide by the numbering */

	if (mode == 0 || mesh->n_nodes < 64) {
		return NULL;
	}

	if (cached != NULL && cached->orig == mesh && cached->n_nodes == mesh->n_nodes && cached->n_elems == mesh->n_elems && cached->kind == (int) mesh->kind && cached->hash == hash) {
		refresh_coords(cached, mesh);
		__atomic_add_fetch(&cached->refs, 1, __ATOMIC_RELAXED);
		return cached;
	}

	if (mode < 0) {
		if (kept.mesh == mesh && kept.n_nodes == mesh->n_nodes && kept.n_elems == mesh->n_elems && kept.hash == hash) {
			return NULL;
		}

		if (mesh->n_nodes <= ((size_t) 1 << 20) || mean_span(mesh) <= (double) ((size_t) 1 << 20)) {
			kept.mesh = mesh, kept.n_nodes = mesh->n_nodes, kept.n_elems = mesh->n_elems, kept.hash = hash;
			return NULL;
		}
	}

	bfmi_renum_t* const r = renum_build(mesh, hash);

	if (r == NULL) {
		return NULL; /* out of memory or bad connectivity: the caller's numbering is used as it is */
	}

	refresh_coords(r, mesh);

	bfmi_renum_release(cached);
	cached = r;

	__atomic_add_fetch(&r->refs, 1, __ATOMIC_RELAXED);
	return r;
}

void bfmi_renum_forget(bfm_mesh_t const* mesh) {
	if (cached != NULL && cached->orig == mesh) {
		bfmi_renum_release(cached);
		cached = NULL;
	}

	if (kept.mesh == mesh) {
		kept.mesh = NULL;
	}
}

/* device copy of to_new (the download gathers x[to_new[a]] into the caller's node a) */
int bfmi_renum_upload(bfmi_renum_t* r) {
	if (r->d_to_new != NULL) {
		return 0;
	}

	if (bfmg_alloc((void**) &r->d_to_new, r->n_nodes * sizeof *r->d_to_new) < 0 || bfmg_upload(r->d_to_new, r->to_new, r->n_nodes * sizeof *r->d_to_new) < 0) {
		bfmg_free(r->d_to_new);
		r->d_to_new = NULL;
		return -1;
	}

	return 0;
}

/* ---- introspection (bfm_b200.h) ------------------------------------------------------------------------------ */

int bfmx_mesh_internal_numbering(bfm_mesh_t* mesh, int32_t* to_new) {
	if (mesh->dim != 2 || mesh->n_nodes == 0 || mesh->n_nodes >= (1u << 30)) {
		return -1;
	}

	bfmi_renum_t* const r = bfmi_renum_for_mesh(mesh, bfmi_mesh_hash(mesh));

	if (r == NULL) {
		return 0;
	}

	if (to_new != NULL) {
		memcpy(to_new, r->to_new, r->n_nodes * sizeof *to_new);
	}

	bfmi_renum_release(r);
	return 1;
}
