/*
 * mesh.c - mesh container, the LEPL1110 and Wavefront readers, edge derivation and the synthetic
 * plate generator used by the benchmarks.
 *
 * Host-side input stage of the hot path ("mesh").  The readers accept the files
 * the reference's fscanf-based readers accept (mesh.c:104-279) and produce the same arrays - edge
 * order included, because Neumann loads are accumulated in edge order (system.c:477-522).
 */
#include "internal.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* reference mesh.c:6-14 */
int bfm_mesh_create(bfm_mesh_t* mesh, bfm_state_t* state, size_t dim, bfm_elem_kind_t kind) {
	memset(mesh, 0, sizeof *mesh);

	mesh->state = state;
	mesh->dim = dim;
	mesh->kind = kind;

	return 0;
}

/* reference mesh.c:16-30 */
int bfm_mesh_destroy(bfm_mesh_t* mesh) {
	bfm_state_t* const state = mesh->state;

	bfmi_plan_forget(mesh); /* drop any cached symbolic plan keyed on this mesh */
	bfmi_part_forget(mesh); /* ... and its row partition */
	bfmi_coarse_forget(mesh); /* ... and the solver's aggregates */
	bfmi_renum_forget(mesh); /* ... and its internally renumbered copy */

	bfmg_host_unpin(mesh->coords); /* bfm_sim_run page-locks large coordinate arrays in place */

	state->free(mesh->coords);
	state->free(mesh->elems);
	state->free(mesh->edges);

	for (size_t i = 0; i < mesh->n_domains; i++) {
		state->free(mesh->domains[i].elements);
	}

	state->free(mesh->domains);
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * edge derivation (reference compute_edges, mesh.c:32-102)
 *
 * One half-edge per element side, sorted by (smaller node DESCENDING, larger node ascending) with a
 * STABLE sort (glibc's qsort is a merge sort), then neighbouring opposite half-edges are fused into
 * one interior edge.  A trailing unpaired half-edge is dropped, exactly like the reference's loop.
 * ------------------------------------------------------------------------------------------- */

static inline bool half_edge_before(bfm_edge_t const* a, bfm_edge_t const* b) {
	/* true when a must come strictly before b; equal keys keep their input order */

	int const a_lo = (int) BFM_MIN(a->nodes[0], a->nodes[1]);
	int const b_lo = (int) BFM_MIN(b->nodes[0], b->nodes[1]);

	if (a_lo != b_lo) {
		return a_lo > b_lo;
	}

	int const a_hi = (int) BFM_MAX(a->nodes[0], a->nodes[1]);
	int const b_hi = (int) BFM_MAX(b->nodes[0], b->nodes[1]);

	return a_hi < b_hi;
}

static void half_edge_sort(bfm_edge_t* v, bfm_edge_t* tmp, size_t n) {
	if (n < 2) {
		return;
	}

	if (n <= 12) {
		for (size_t i = 1; i < n; i++) {
			bfm_edge_t const cur = v[i];
			size_t j = i;

			for (; j > 0 && half_edge_before(&cur, &v[j - 1]); j--) {
				v[j] = v[j - 1];
			}

			v[j] = cur;
		}

		return;
	}

	size_t const half = n / 2;

	/* the two halves use disjoint scratch, so big ones can be sorted by different threads (OpenMP tasks, when
	 * called from a parallel region); the merge keeps the result independent of who sorted what */

#pragma omp task shared(v, tmp) if (n > ((size_t) 1 << 18))
	half_edge_sort(v, tmp, half);

#pragma omp task shared(v, tmp) if (n > ((size_t) 1 << 18))
	half_edge_sort(v + half, tmp + half, n - half);

#pragma omp taskwait

	size_t l = 0, r = half, o = 0;

	while (l < half && r < n) {
		tmp[o++] = half_edge_before(&v[r], &v[l]) ? v[r++] : v[l++];
	}

	while (l < half) {
		tmp[o++] = v[l++];
	}

	while (r < n) {
		tmp[o++] = v[r++];
	}

	memcpy(v, tmp, n * sizeof *v);
}

/* the same list from the kernels of symbolic.cu (per-node segments instead of a global sort; identical order).
 * 1: done, 0: not applicable here (the caller takes the host path), -1: failed */
static int edges_on_device(bfm_mesh_t* mesh) {
	bfm_state_t* const state = mesh->state;
	size_t const sides = mesh->kind;
	size_t const n_half = mesh->n_elems * sides;
	size_t const nn = mesh->n_nodes;

	if ((sides != 3 && sides != 4) || nn == 0 || nn >= (1u << 30) || n_half >= INT32_MAX || n_half < 2) {
		return 0;
	}

	int32_t* d_elems = NULL;
	int64_t* d_edges = NULL;
	int64_t n_out = 0;
	int bad = 0;
	int rv = -1;

	/* connectivity narrowed to 32 bits on its way to the device; node numbers outside the table: the host sort does
	 * not mind them, the per-node segments would - the caller takes the host path */
	if (bfmg_alloc((void**) &d_elems, n_half * sizeof *d_elems) < 0 || bfmg_upload_narrow(d_elems, mesh->elems, n_half, nn, &bad) < 0) {
		BFMI_FAIL(state, "edge derivation on the device failed: %s", bfmg_last_error());
		bfmg_free(d_elems);
		return -1;
	}

	if (bad) {
		bfmg_free(d_elems);
		return 0;
	}

	if (bfmg_edges_build((int32_t) nn, (int64_t) mesh->n_elems, (int32_t) sides, d_elems, &d_edges, &n_out) == 0) {
		_Static_assert(sizeof(bfm_edge_t) == 4 * sizeof(int64_t), "an edge record is nodes[2], elems[2], 8 bytes each");

		mesh->edges = n_out > 0 ? state->alloc((size_t) n_out * sizeof *mesh->edges) : NULL;

		if (mesh->edges != NULL && bfmg_download(mesh->edges, d_edges, (size_t) n_out * sizeof *mesh->edges) == 0) {
			mesh->n_edges = (size_t) n_out;
			rv = 1;
		}

		else if (mesh->edges != NULL) {
			state->free(mesh->edges);
			mesh->edges = NULL;
		}
	}

	else {
		BFMI_FAIL(state, "edge derivation on the device failed: %s", bfmg_last_error());
	}

	bfmg_free(d_elems);
	bfmg_free(d_edges);

	return rv;
}

int bfmx_mesh_compute_edges(bfm_mesh_t* mesh) {
	bfm_state_t* const state = mesh->state;
	size_t const sides = mesh->kind;
	size_t const n_half = mesh->n_elems * sides;

	if (mesh->edges != NULL) {
		state->free(mesh->edges);
		mesh->edges = NULL;
		mesh->n_edges = 0;
	}

	/* big meshes go through the device when there is one (BFM_EDGES=device / host forces either way; small meshes
	 * are not worth a CUDA context: the readers must stay usable on a machine without a GPU) */

	char const* const how = getenv("BFM_EDGES");
	bool const want_device = how != NULL ? strcmp(how, "device") == 0 : n_half >= ((size_t) 1 << 20);

	if (want_device && bfmg_available()) {
		int const rv = edges_on_device(mesh);

		if (rv != 0) {
			return rv < 0 ? -1 : 0;
		}
	}

	bfm_edge_t* const edges = state->alloc(n_half * sizeof *edges);
	bfm_edge_t* const tmp = malloc((n_half + 1) * sizeof *tmp);

	if (edges == NULL || tmp == NULL) {
		free(tmp);

		if (edges != NULL) {
			state->free(edges);
		}

		return -1;
	}

#pragma omp parallel for schedule(static) if (n_half > ((size_t) 1 << 18))
	for (size_t e = 0; e < mesh->n_elems; e++) {
		for (size_t j = 0; j < sides; j++) {
			bfm_edge_t* const he = &edges[e * sides + j];

			he->nodes[0] = mesh->elems[e * sides + j];
			he->nodes[1] = mesh->elems[e * sides + (j + 1) % sides];
			he->elems[0] = (ssize_t) e;
			he->elems[1] = -1;
		}
	}

#pragma omp parallel if (n_half > ((size_t) 1 << 18))
#pragma omp single
	half_edge_sort(edges, tmp, n_half);

	free(tmp);

	size_t n_out = 0;

	for (size_t i = 0; i + 1 < n_half; i++) {
		bfm_edge_t const first = edges[i];
		bfm_edge_t const next = edges[i + 1];

		edges[n_out] = first;

		if (first.nodes[0] == next.nodes[1] && first.nodes[1] == next.nodes[0]) {
			edges[n_out].elems[1] = next.elems[0];
			i++; /* the partner is consumed */
		}

		n_out++;
	}

	mesh->n_edges = n_out;
	mesh->edges = state->realloc(edges, n_out * sizeof *edges);

	return mesh->edges == NULL ? -1 : 0; /* realloc(.., 0) -> NULL -> failure, like the reference */
}

/* ---------------------------------------------------------------------------------------------
 * a tiny scanner over an in-memory copy of the file: the same tokens fscanf would see
 * ------------------------------------------------------------------------------------------- */

typedef struct {
	char* buf;
	char const* at;
	char const* end; /* the terminating NUL */
} scan_t;

static int scan_open(scan_t* sc, char const* path) {
	FILE* const fp = fopen(path, "rb");

	if (fp == NULL) {
		return -1;
	}

	fseek(fp, 0, SEEK_END);
	long const size = ftell(fp);
	fseek(fp, 0, SEEK_SET);

	sc->buf = malloc((size_t) size + 1);

	if (sc->buf == NULL || fread(sc->buf, 1, (size_t) size, fp) != (size_t) size) {
		fclose(fp);
		free(sc->buf);
		return -1;
	}

	fclose(fp);

	sc->buf[size] = '\0';
	sc->at = sc->buf;
	sc->end = sc->buf + size;

	return 0;
}

static void scan_ws(scan_t* sc) {
	while (*sc->at != '\0' && isspace((unsigned char) *sc->at)) {
		sc->at++;
	}
}

/* match a literal; blanks in the literal stand for "any amount of whitespace" (as in scanf) */
static bool scan_lit(scan_t* sc, char const* lit) {
	char const* const start = sc->at;

	for (; *lit != '\0'; lit++) {
		if (*lit == ' ') {
			scan_ws(sc);
		}

		else if (*sc->at == *lit) {
			sc->at++;
		}

		else {
			sc->at = start;
			return false;
		}
	}

	return true;
}

static bool scan_size(scan_t* sc, size_t* out) {
	scan_ws(sc);

	if (!isdigit((unsigned char) *sc->at) && *sc->at != '-' && *sc->at != '+') {
		return false;
	}

	char* end;
	*out = (size_t) strtoull(sc->at, &end, 10);

	if (end == sc->at) {
		return false;
	}

	sc->at = end;
	return true;
}

static bool scan_double(scan_t* sc, double* out) {
	scan_ws(sc);

	char* end;
	*out = strtod(sc->at, &end);

	if (end == sc->at) {
		return false;
	}

	sc->at = end;
	return true;
}

/* next blank-delimited token, at most max - 1 characters (scanf's %Ns) */
static bool scan_word(scan_t* sc, char* out, size_t max) {
	scan_ws(sc);

	size_t len = 0;

	while (*sc->at != '\0' && !isspace((unsigned char) *sc->at) && len + 1 < max) {
		out[len++] = *sc->at++;
	}

	out[len] = '\0';
	return len > 0;
}

/* rest of the current line without its newline (scanf's %[^\n]), at most max - 1 characters */
static bool scan_rest(scan_t* sc, char* out, size_t max) {
	size_t len = 0;

	while (*sc->at != '\0' && *sc->at != '\n') {
		if (len + 1 < max) {
			out[len++] = *sc->at;
		}

		sc->at++;
	}

	out[len] = '\0';
	return len > 0;
}

static void scan_skip_line(scan_t* sc) {
	while (*sc->at != '\0' && *sc->at++ != '\n') {
	}
}

/* ---------------------------------------------------------------------------------------------
 * big files: the long sections, one record per line, parsed by all threads
 *
 * A 25 M-node mesh is gigabytes of text, and a single thread converting it with strtod takes the better part of a
 * minute.  The sections that matter ("i : x y" nodes and "i : a b c [d]" elements of a LEPL1110 file, "v" and "f"
 * lines of an OBJ) are written one record per line by every tool around, so: find the line starts in parallel
 * (newlines counted per 1 MB chunk, prefix sum), let every thread parse its share of the lines with the SAME token
 * scanners as the serial reader, and check that each line held exactly one record and nothing else.  scanf does not
 * care about line breaks, so a file may legally split or join records differently; the first line that is not
 * exactly one record makes the whole section fall back to the serial scanner (whose behaviour is the reference's
 * fscanf's), so the fast path can only ever produce what the serial one would.  BFM_READER=serial switches it off.
 * ------------------------------------------------------------------------------------------- */

#define PAR_CHUNK ((size_t) 1 << 20)
#define PAR_MIN_BYTES ((size_t) 1 << 20) /* smaller inputs are not worth the threads */

static bool reader_parallel(size_t bytes) {
	char const* const env = getenv("BFM_READER");
	return bytes >= PAR_MIN_BYTES && (env == NULL || strcmp(env, "serial") != 0);
}

/* line starts from `from` on: starts[0] = from, starts[i] = one past the i-th newline; at most `want` lines.
 * Returns the lines found (each line i is [starts[i], starts[i + 1])), or 0 without memory; *out is malloc'd */
static size_t index_lines(char const* from, char const* end, size_t want, char const*** out) {
	size_t const bytes = (size_t) (end - from);
	size_t const n_chunks = (bytes + PAR_CHUNK - 1) / PAR_CHUNK;
	size_t* const count = calloc(n_chunks + 2, sizeof *count);

	*out = NULL;

	if (count == NULL || bytes == 0 || want == 0) {
		free(count);
		return 0;
	}

#pragma omp parallel for schedule(static)
	for (size_t c = 0; c < n_chunks; c++) {
		char const* at = from + c * PAR_CHUNK;
		char const* const stop = c + 1 < n_chunks ? at + PAR_CHUNK : end;
		size_t n = 0;

		while ((at = memchr(at, '\n', (size_t) (stop - at))) != NULL) {
			n++, at++;
		}

		count[c + 1] = n;
	}

	for (size_t c = 0; c < n_chunks; c++) {
		count[c + 1] += count[c];
	}

	size_t const newlines = count[n_chunks];
	size_t lines = newlines + (end[-1] != '\n'); /* a last line without its newline still counts */

	lines = lines < want ? lines : want;

	char const** const starts = malloc((lines + 2) * sizeof *starts);

	if (starts == NULL) {
		free(count);
		return 0;
	}

	starts[0] = from;
	starts[lines] = end; /* overwritten below when the file goes on after the last wanted line */

#pragma omp parallel for schedule(static)
	for (size_t c = 0; c < n_chunks; c++) {
		char const* at = from + c * PAR_CHUNK;
		char const* const stop = c + 1 < n_chunks ? at + PAR_CHUNK : end;
		size_t idx = count[c];

		while (idx < lines && (at = memchr(at, '\n', (size_t) (stop - at))) != NULL) {
			starts[++idx] = ++at;
		}
	}

	free(count);

	*out = starts;
	return lines;
}

/* only blanks between at and stop? (and the scanner did not run past the line) */
static bool line_done(char const* at, char const* stop) {
	if (at > stop) {
		return false;
	}

	for (; at < stop; at++) {
		if (!isspace((unsigned char) *at)) {
			return false;
		}
	}

	return true;
}

/* `n` records "index : v v .." of `width` values each (doubles into dvals, else sizes into svals), one per line, from
 * the scanner's position on.  true: all parsed, scanner moved behind them.  false: not in that shape (or too small to
 * bother): nothing consumed, the caller scans them one by one */
static bool scan_records_parallel(scan_t* sc, size_t n, size_t width, double* dvals, size_t* svals) {
	if (n < 4096 || !reader_parallel((size_t) (sc->end - sc->at))) {
		return false;
	}

	scan_t first = *sc;

	scan_ws(&first); /* the rest of the header line */

	char const** starts;
	size_t const lines = index_lines(first.at, sc->end, n, &starts);

	if (lines != n) {
		free(starts);
		return false;
	}

	bool ok = true;

#pragma omp parallel for schedule(static) reduction(&& : ok)
	for (size_t i = 0; i < n; i++) {
		scan_t line = {.buf = NULL, .at = starts[i], .end = sc->end};
		size_t index;
		bool good = scan_size(&line, &index) && scan_lit(&line, " :");

		for (size_t k = 0; good && k < width; k++) {
			good = dvals != NULL ? scan_double(&line, &dvals[i * width + k]) : scan_size(&line, &svals[i * width + k]);
		}

		ok = ok && good && line_done(line.at, starts[i + 1]);
	}

	if (ok) {
		sc->at = starts[n];
	}

	free(starts);
	return ok;
}

/* ---------------------------------------------------------------------------------------------
 * LEPL1110 text meshes (reference mesh.c:104-201)
 *
 *   Number of nodes N          then N lines  "i : x y"
 *   Number of edges N          then N lines  "elem : n0 n1"       (boundary edges only)
 *   Number of triangles|quads  then N lines  "i : a b c [d]"
 *   Number of domains N        then per domain "Domain : id / Name : .. / Number of elements : n" + ids
 *
 * All four sections are mandatory: files without the edge/domain sections (meshes/gear*.lepl1110)
 * are rejected with -1, as by the reference.
 * ------------------------------------------------------------------------------------------- */

int bfm_mesh_read_lepl1110(bfm_mesh_t* mesh, bfm_state_t* state, char const* name) {
	memset(mesh, 0, sizeof *mesh);

	mesh->state = state;
	mesh->dim = 2;

	scan_t sc;

	if (scan_open(&sc, name) < 0) {
		return -1;
	}

	int rv = -1;
	size_t index;

	/* nodes */

	if (!scan_lit(&sc, " Number of nodes") || !scan_size(&sc, &mesh->n_nodes)) {
		goto done;
	}

	mesh->coords = state->alloc(mesh->n_nodes * 2 * sizeof *mesh->coords);

	if (mesh->coords == NULL) {
		goto done;
	}

	for (size_t i = scan_records_parallel(&sc, mesh->n_nodes, 2, mesh->coords, NULL) ? mesh->n_nodes : 0; i < mesh->n_nodes; i++) {
		if (!scan_size(&sc, &index) || !scan_lit(&sc, " :") || !scan_double(&sc, &mesh->coords[2 * i]) || !scan_double(&sc, &mesh->coords[2 * i + 1])) {
			goto done;
		}
	}

	/* boundary edges */

	if (!scan_lit(&sc, " Number of edges") || !scan_size(&sc, &mesh->n_edges)) {
		goto done;
	}

	mesh->edges = state->alloc(mesh->n_edges * sizeof *mesh->edges);

	if (mesh->edges == NULL) {
		goto done;
	}

	for (size_t i = 0; i < mesh->n_edges; i++) {
		bfm_edge_t* const edge = &mesh->edges[i];
		size_t elem;

		if (!scan_size(&sc, &elem) || !scan_lit(&sc, " :") || !scan_size(&sc, &edge->nodes[0]) || !scan_size(&sc, &edge->nodes[1])) {
			goto done;
		}

		edge->elems[0] = (ssize_t) elem;
		edge->elems[1] = -1;
	}

	/* elements */

	char kind_str[16];

	if (!scan_lit(&sc, " Number of") || !scan_word(&sc, kind_str, sizeof kind_str) || !scan_size(&sc, &mesh->n_elems)) {
		goto done;
	}

	if (strcmp(kind_str, "triangles") == 0) {
		mesh->kind = BFM_ELEM_KIND_SIMPLEX;
	}

	else if (strcmp(kind_str, "quads") == 0) {
		mesh->kind = BFM_ELEM_KIND_QUAD;
	}

	else {
		goto done;
	}

	mesh->elems = state->alloc(mesh->n_elems * mesh->kind * sizeof *mesh->elems);

	if (mesh->elems == NULL) {
		goto done;
	}

	for (size_t i = scan_records_parallel(&sc, mesh->n_elems, mesh->kind, NULL, mesh->elems) ? mesh->n_elems : 0; i < mesh->n_elems; i++) {
		if (!scan_size(&sc, &index) || !scan_lit(&sc, " :")) {
			goto done;
		}

		for (size_t j = 0; j < (size_t) mesh->kind; j++) {
			if (!scan_size(&sc, &mesh->elems[i * mesh->kind + j])) {
				goto done;
			}
		}
	}

	/* domains */

	if (!scan_lit(&sc, " Number of domains") || !scan_size(&sc, &mesh->n_domains)) {
		goto done;
	}

	mesh->domains = state->alloc(mesh->n_domains * sizeof *mesh->domains);

	if (mesh->domains == NULL) {
		goto done;
	}

	memset(mesh->domains, 0, mesh->n_domains * sizeof *mesh->domains);

	for (size_t i = 0; i < mesh->n_domains; i++) {
		size_t id;

		if (!scan_lit(&sc, " Domain :") || !scan_size(&sc, &id) || id >= mesh->n_domains) {
			goto done;
		}

		bfm_domain_t* const domain = &mesh->domains[id]; /* stored under its id (mesh.c:167) */

		if (!scan_lit(&sc, " Name : ")) {
			goto done;
		}

		scan_rest(&sc, domain->name, sizeof domain->name); /* keeps trailing blanks ("Entity 1 ") */

		if (!scan_lit(&sc, " Number of elements :") || !scan_size(&sc, &domain->n_elements)) {
			goto done;
		}

		if (domain->elements != NULL) {
			state->free(domain->elements);
		}

		domain->elements = state->alloc(domain->n_elements * sizeof *domain->elements);

		if (domain->elements == NULL) {
			goto done;
		}

		for (size_t j = 0; j < domain->n_elements; j++) {
			if (!scan_size(&sc, &domain->elements[j])) {
				goto done;
			}
		}
	}

	rv = 0;

done:

	free(sc.buf);
	return rv; /* on failure the partially filled mesh is left to bfm_mesh_destroy, as in the reference */
}

/* ---------------------------------------------------------------------------------------------
 * Wavefront OBJ (reference mesh.c:203-279): "o name", "v x y z", "f a b c" (1-based); every other
 * line is skipped.  z is kept only when `full`.
 * ------------------------------------------------------------------------------------------- */

/* what one line of an OBJ file is to the reader */
enum { OBJ_SKIP, OBJ_VERTEX, OBJ_FACE, OBJ_ODD };

/* classifies the line [at, stop) and, when out pointers are given, parses it.  OBJ_ODD: a line the serial scanner
 * would not treat as one self-contained record (a record running over the line end, or something after it that could
 * itself be a record) */
static int obj_line(char const* at, char const* stop, char const* end, size_t dim, double* xyz_out, size_t* tri_out) {
	scan_t line = {.buf = NULL, .at = at, .end = end};
	char header[16];

	scan_ws(&line);

	if (line.at >= stop) {
		return OBJ_SKIP; /* blank */
	}

	scan_word(&line, header, sizeof header);

	if (strcmp(header, "v") == 0) {
		double xyz[3];

		for (size_t k = 0; k < 3; k++) {
			if (!scan_double(&line, &xyz[k]) || line.at > stop) {
				return OBJ_ODD;
			}
		}

		if (!line_done(line.at, stop)) {
			return OBJ_ODD;
		}

		if (xyz_out != NULL) {
			memcpy(xyz_out, xyz, dim * sizeof *xyz);
		}

		return OBJ_VERTEX;
	}

	if (strcmp(header, "f") == 0) {
		size_t tri[3];

		for (size_t k = 0; k < 3; k++) {
			if (!scan_size(&line, &tri[k]) || line.at > stop) {
				return OBJ_ODD;
			}

			tri[k]--;

			while (*line.at != '\0' && !isspace((unsigned char) *line.at)) { /* "a/at/an" */
				line.at++;
			}
		}

		if (!line_done(line.at, stop)) {
			return OBJ_ODD;
		}

		if (tri_out != NULL) {
			memcpy(tri_out, tri, sizeof tri);
		}

		return OBJ_FACE;
	}

	if (strcmp(header, "o") == 0) {
		char word[256];

		/* "o name": the serial scanner takes the next word as the name, and whatever follows on the line as new
		 * header words */
		if (!scan_word(&line, word, sizeof word) || line.at > stop) {
			return OBJ_ODD;
		}

		return line_done(line.at, stop) ? OBJ_SKIP : OBJ_ODD;
	}

	/* anything else: the serial scanner skips the rest of the line; a header of 15+ characters is cut by %15s and its
	 * tail would be read as the next header - leave that to it */
	return strlen(header) < sizeof header - 1 ? OBJ_SKIP : OBJ_ODD;
}

/* the whole file by lines, in parallel (see "big files" above).  1: mesh->coords / elems filled, 0: not in that shape,
 * -1: out of memory */
static int read_wavefront_parallel(bfm_mesh_t* mesh, scan_t* sc) {
	bfm_state_t* const state = mesh->state;
	size_t const bytes = (size_t) (sc->end - sc->buf);

	if (!reader_parallel(bytes)) {
		return 0;
	}

	char const** starts;
	size_t const lines = index_lines(sc->buf, sc->end, SIZE_MAX / 16, &starts);

	if (lines == 0) {
		return 0;
	}

	size_t const block = 4096; /* lines per counting block */
	size_t const n_blocks = (lines + block - 1) / block;
	size_t* const first_v = calloc(n_blocks + 1, sizeof *first_v);
	size_t* const first_f = calloc(n_blocks + 1, sizeof *first_f);
	bool ok = first_v != NULL && first_f != NULL;
	int rv = 0;

	if (!ok) {
		rv = -1;
		goto out;
	}

#pragma omp parallel for schedule(static) reduction(&& : ok)
	for (size_t b = 0; b < n_blocks; b++) {
		size_t const stop = (b + 1) * block < lines ? (b + 1) * block : lines;

		for (size_t i = b * block; i < stop; i++) {
			int const what = obj_line(starts[i], starts[i + 1], sc->end, mesh->dim, NULL, NULL);

			first_v[b + 1] += what == OBJ_VERTEX;
			first_f[b + 1] += what == OBJ_FACE;
			ok = ok && what != OBJ_ODD;
		}
	}

	if (!ok) {
		goto out;
	}

	for (size_t b = 0; b < n_blocks; b++) {
		first_v[b + 1] += first_v[b];
		first_f[b + 1] += first_f[b];
	}

	mesh->n_nodes = first_v[n_blocks];
	mesh->n_elems = first_f[n_blocks];
	mesh->coords = mesh->n_nodes > 0 ? state->alloc(mesh->n_nodes * mesh->dim * sizeof *mesh->coords) : NULL;
	mesh->elems = mesh->n_elems > 0 ? state->alloc(mesh->n_elems * 3 * sizeof *mesh->elems) : NULL;

	if ((mesh->n_nodes > 0 && mesh->coords == NULL) || (mesh->n_elems > 0 && mesh->elems == NULL)) {
		state->free(mesh->coords);
		state->free(mesh->elems);
		mesh->coords = NULL, mesh->elems = NULL;
		mesh->n_nodes = mesh->n_elems = 0;
		rv = -1;
		goto out;
	}

#pragma omp parallel for schedule(static)
	for (size_t b = 0; b < n_blocks; b++) {
		size_t const stop = (b + 1) * block < lines ? (b + 1) * block : lines;
		size_t v = first_v[b], f = first_f[b];

		for (size_t i = b * block; i < stop; i++) {
			int const what = obj_line(starts[i], starts[i + 1], sc->end, mesh->dim, mesh->coords != NULL ? &mesh->coords[v * mesh->dim] : NULL, mesh->elems != NULL ? &mesh->elems[f * 3] : NULL);

			v += what == OBJ_VERTEX;
			f += what == OBJ_FACE;
		}
	}

	rv = 1;

out:

	free(first_v);
	free(first_f);
	free(starts);

	return rv;
}

int bfm_mesh_read_wavefront(bfm_mesh_t* mesh, bfm_state_t* state, char const* name, bool full) {
	memset(mesh, 0, sizeof *mesh);

	mesh->state = state;
	mesh->dim = full ? 3 : 2;
	mesh->kind = BFM_ELEM_KIND_SIMPLEX;

	scan_t sc;

	if (scan_open(&sc, name) < 0) {
		return -1;
	}

	size_t cap_nodes = 0;
	size_t cap_elems = 0;
	char header[16];
	int rv = -1;

	int const fast = read_wavefront_parallel(mesh, &sc);

	if (fast < 0) {
		goto done;
	}

	while (fast == 0 && scan_word(&sc, header, sizeof header)) {
		if (strcmp(header, "o") == 0) {
			char obj_name[256];
			scan_word(&sc, obj_name, sizeof obj_name);
		}

		else if (strcmp(header, "v") == 0) {
			if (mesh->n_nodes == cap_nodes) {
				cap_nodes = cap_nodes ? 2 * cap_nodes : 1024;
				mesh->coords = state->realloc(mesh->coords, cap_nodes * mesh->dim * sizeof *mesh->coords);

				if (mesh->coords == NULL) {
					goto done;
				}
			}

			double xyz[3] = {0, 0, 0};

			for (size_t k = 0; k < 3 && scan_double(&sc, &xyz[k]); k++) {
			}

			memcpy(&mesh->coords[mesh->n_nodes++ * mesh->dim], xyz, mesh->dim * sizeof *xyz);
		}

		else if (strcmp(header, "f") == 0) {
			if (mesh->n_elems == cap_elems) {
				cap_elems = cap_elems ? 2 * cap_elems : 1024;
				mesh->elems = state->realloc(mesh->elems, cap_elems * 3 * sizeof *mesh->elems);

				if (mesh->elems == NULL) {
					goto done;
				}
			}

			size_t* const tri = &mesh->elems[mesh->n_elems++ * 3];

			for (size_t k = 0; k < 3; k++) {
				if (!scan_size(&sc, &tri[k])) {
					goto done;
				}

				tri[k]--;

				/* tolerate "a/at/an" triplets: only the vertex index matters */

				while (*sc.at != '\0' && !isspace((unsigned char) *sc.at)) {
					sc.at++;
				}
			}
		}

		else {
			scan_skip_line(&sc);
		}
	}

	if (bfmx_mesh_compute_edges(mesh) < 0) {
		/* the reference releases the arrays here and still reports failure (mesh.c:262-267) */

		state->free(mesh->coords);
		state->free(mesh->elems);
		mesh->coords = NULL;
		mesh->elems = NULL;

		goto done;
	}

	rv = 0;

done:

	free(sc.buf);
	return rv;
}

/* ---------------------------------------------------------------------------------------------
 * synthetic structured plate: [0,lx] x [0,ly],
 * (nx+1)(ny+1) nodes numbered row by row, node (i,j) at (lx*i/nx, ly*j/ny); each cell
 * a=(i,j), b=a+1, c=a+nx+1, d=c+1 becomes triangles (a,b,d),(a,d,c) or the quad (d,c,a,b).
 * Only rows [j0, j1] of nodes are generated when a strip is requested (multi-GPU partitions build
 * just their own part); pass j0 = 0, j1 = ny for the whole plate.
 * ------------------------------------------------------------------------------------------- */

int bfmx_mesh_plate(bfm_mesh_t* mesh, bfm_state_t* state, size_t nx, size_t ny, double lx, double ly, bfm_elem_kind_t kind, bool with_edges) {
	if (nx == 0 || ny == 0 || (kind != BFM_ELEM_KIND_SIMPLEX && kind != BFM_ELEM_KIND_QUAD)) {
		return -1;
	}

	bfm_mesh_create(mesh, state, 2, kind);

	size_t const per_cell = kind == BFM_ELEM_KIND_SIMPLEX ? 2 : 1;

	mesh->n_nodes = (nx + 1) * (ny + 1);
	mesh->n_elems = per_cell * nx * ny;

	mesh->coords = state->alloc(mesh->n_nodes * 2 * sizeof *mesh->coords);
	mesh->elems = state->alloc(mesh->n_elems * kind * sizeof *mesh->elems);

	if (mesh->coords == NULL || mesh->elems == NULL) {
		return -1;
	}

#pragma omp parallel for schedule(static) if (nx * ny > 100000)
	for (size_t j = 0; j <= ny; j++) {
		for (size_t i = 0; i <= nx; i++) {
			size_t const node = j * (nx + 1) + i;

			mesh->coords[2 * node + 0] = lx * (double) i / (double) nx;
			mesh->coords[2 * node + 1] = ly * (double) j / (double) ny;
		}
	}

#pragma omp parallel for schedule(static) if (nx * ny > 100000)
	for (size_t j = 0; j < ny; j++) {
		for (size_t i = 0; i < nx; i++) {
			size_t const a = j * (nx + 1) + i;
			size_t const b = a + 1;
			size_t const c = a + nx + 1;
			size_t const d = c + 1;
			size_t* const out = &mesh->elems[(j * nx + i) * per_cell * kind];

			if (kind == BFM_ELEM_KIND_SIMPLEX) {
				out[0] = a, out[1] = b, out[2] = d;
				out[3] = a, out[4] = d, out[5] = c;
			}

			else {
				out[0] = d, out[1] = c, out[2] = a, out[3] = b;
			}
		}
	}

	return with_edges ? bfmx_mesh_compute_edges(mesh) : 0;
}
