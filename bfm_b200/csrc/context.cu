/*
 * context.cu - device context of libbfm (B200 build): one device, one non-blocking stream, a
 * stream-ordered memory pool, an event stopwatch and the kernel-launch counter.
 *
 * Part of the thin C ABI declared in gpu.h.  If no CUDA device can be initialised every entry point
 * fails with -1 and a message - the library has no CPU path.
 */
#include "gpu_internal.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace {

struct Context {
	bool tried = false;
	bool ok = false;
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	size_t launches = 0;
	char err[512] = "";

	static constexpr int kEvents = 64;
	cudaEvent_t events[kEvents];
	int next_event = 0;

	cudaEvent_t user[8][2] = {}; /* bfmx_timer_* slots */

	void* pinned = nullptr;      /* 4 KiB of page-locked host memory for status polls */
	cudaEvent_t poll[2] = {};    /* untimed events marking those polls */

	/* large host <-> device copies go through two page-locked staging buffers (allocated on first use) */
	char* stage[2] = {};
	cudaEvent_t stage_done[2] = {};
	bool stage_busy[2] = {};

	/* caller buffers page-locked in place (bfmg_host_pin): [base, base + bytes) */
	struct Pin { char* base; size_t bytes; } pins[8] = {};
};

constexpr size_t kStageBytes = (size_t) 32 << 20;
constexpr size_t kStageFrom = (size_t) 4 << 20; /* smaller copies are left to the driver */

Context G;

void set_error(char const* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(G.err, sizeof G.err, fmt, ap);
	va_end(ap);
}

bool init() {
	if (G.tried) {
		return G.ok;
	}

	G.tried = true;

	int count = 0;
	cudaError_t rc = cudaGetDeviceCount(&count);

	if (rc != cudaSuccess || count == 0) {
		set_error("no usable CUDA device (%s); libbfm's hot path is GPU-only and has no CPU fallback", rc != cudaSuccess ? cudaGetErrorString(rc) : "device count is 0");
		cudaGetLastError();
		return false;
	}

	/* one process per GPU: honour BFM_DEVICE, else LOCAL_RANK (torchrun), else the current device */

	char const* env = getenv("BFM_DEVICE");

	if (env == nullptr) {
		env = getenv("LOCAL_RANK");
	}

	if (env != nullptr) {
		G.device = atoi(env) % count;
	}

	else if (cudaGetDevice(&G.device) != cudaSuccess) {
		G.device = 0;
	}

	if ((rc = cudaSetDevice(G.device)) != cudaSuccess) {
		set_error("cudaSetDevice(%d): %s", G.device, cudaGetErrorString(rc));
		return false;
	}

	cudaDeviceProp prop;

	if ((rc = cudaGetDeviceProperties(&prop, G.device)) != cudaSuccess) {
		set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(rc));
		return false;
	}

	G.sm_count = prop.multiProcessorCount;

	if ((rc = cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking)) != cudaSuccess) {
		set_error("cudaStreamCreate: %s", cudaGetErrorString(rc));
		return false;
	}

	/* keep freed blocks in the pool: repeated bfm_sim_run calls reuse them without going to the driver */

	cudaMemPool_t pool;

	if (cudaDeviceGetDefaultMemPool(&pool, G.device) == cudaSuccess) {
		uint64_t keep = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	}

	for (auto& ev : G.events) {
		if ((rc = cudaEventCreate(&ev)) != cudaSuccess) {
			set_error("cudaEventCreate: %s", cudaGetErrorString(rc));
			return false;
		}
	}

	for (auto& pair : G.user) {
		for (auto& ev : pair) {
			if ((rc = cudaEventCreate(&ev)) != cudaSuccess) {
				set_error("cudaEventCreate: %s", cudaGetErrorString(rc));
				return false;
			}
		}
	}

	for (auto& ev : G.poll) {
		if ((rc = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) {
			set_error("cudaEventCreate: %s", cudaGetErrorString(rc));
			return false;
		}
	}

	if ((rc = cudaMallocHost(&G.pinned, 4096)) != cudaSuccess) {
		set_error("cudaMallocHost: %s", cudaGetErrorString(rc));
		return false;
	}

	G.ok = true;
	return true;
}

} // namespace

/* ---- internal accessors (gpu_internal.cuh) ----------------------------------------------------- */

bool bfmg_ready() {
	if (!init()) {
		return false;
	}

	cudaSetDevice(G.device); /* cheap; keeps us correct if the host program switched devices */
	return true;
}

cudaStream_t bfmg_stream() {
	return G.stream;
}

bool bfmg_pdl() {
	static int enabled = -1;

	if (enabled < 0) {
		char const* const env = getenv("BFM_PDL");
		enabled = env != nullptr && atoi(env) != 0 ? 1 : 0; /* off unless asked for: measured, it buys nothing inside a replayed graph */
	}

	return enabled == 1;
}

void* bfmg_pinned() {
	return G.pinned;
}

cudaEvent_t bfmg_poll_event(int i) {
	return G.poll[i & 1];
}

void bfmg_count_launch(size_t n) {
	G.launches += n;
}

void bfmg_set_error(char const* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(G.err, sizeof G.err, fmt, ap);
	va_end(ap);
}

int bfmg_device() {
	return G.device;
}

int bfmg_check(cudaError_t rc, char const* what, char const* file, int line) {
	if (rc == cudaSuccess) {
		return 0;
	}

	set_error("%s: %s (%s:%d)", what, cudaGetErrorString(rc), file, line);
	cudaGetLastError();
	return -1;
}

/* ---- gpu.h --------------------------------------------------------------------------------------- */

extern "C" {

int bfmg_available(void) {
	return bfmg_ready() ? 1 : 0;
}

char const* bfmg_last_error(void) {
	return G.err;
}

int bfmg_sm_count(void) {
	return bfmg_ready() ? G.sm_count : 0;
}

size_t bfmg_launch_count(void) {
	return G.launches;
}

int bfmg_alloc(void** d_ptr, size_t bytes) {
	*d_ptr = nullptr;

	if (!bfmg_ready()) {
		return -1;
	}

	return BFMG_CHECK(cudaMallocAsync(d_ptr, bytes ? bytes : 16, G.stream));
}

void bfmg_free(void* d_ptr) {
	if (d_ptr != nullptr && G.ok) {
		cudaSetDevice(G.device);
		cudaFreeAsync(d_ptr, G.stream);
	}
}

/* The caller's buffers (mesh->coords, instance->effects: whatever state->alloc returned) are pageable, and the
 * driver stages a pageable copy through its own small bounce buffer at ~6 GB/s (measured: 136 ms for the 800 MB
 * of a 50 M-DOF step).  Large copies therefore go through two page-locked 32 MB buffers of the library's own:
 * the host threads copy chunk i + 1 between caller memory and one buffer while the DMA engine moves chunk i
 * between the other and the device - PCIe rate, no pinning of caller memory (cudaHostRegister of 400 MB costs
 * as much as the copy). */
static bool stage_ready() {
	if (G.stage[0] != nullptr) {
		return true;
	}

	for (int i = 0; i < 2; i++) {
		if (cudaMallocHost((void**) &G.stage[i], kStageBytes) != cudaSuccess || cudaEventCreateWithFlags(&G.stage_done[i], cudaEventDisableTiming) != cudaSuccess) {
			cudaGetLastError();

			if (G.stage[0] != nullptr) {
				cudaFreeHost(G.stage[0]);
				G.stage[0] = nullptr;
			}

			return false; /* no staging: the plain path still works */
		}
	}

	return true;
}

static void host_copy(char* dst, char const* src, size_t bytes) {
	size_t const piece = (size_t) 1 << 20;
	long const n = (long) ((bytes + piece - 1) / piece);

#pragma omp parallel for schedule(static) num_threads(8) if (n >= 8)
	for (long i = 0; i < n; i++) {
		size_t const off = (size_t) i * piece;
		memcpy(dst + off, src + off, off + piece <= bytes ? piece : bytes - off);
	}
}

static bool is_pinned(void const* ptr, size_t bytes) {
	char const* const p = (char const*) ptr;

	for (auto const& pin : G.pins) {
		if (pin.base != nullptr && p >= pin.base && p + bytes <= pin.base + pin.bytes) {
			return true;
		}
	}

	return false;
}

/* Buffers that live across calls and are copied whole every time - mesh->coords on the way in, instance->effects on
 * the way out (examples/benchmark.py runs one simulation again and again) - are page-locked in place once, so that the
 * DMA engine reads and writes them directly at PCIe rate with no host thread in the way; on an 8-GPU box, where eight
 * processes each fetch the complete 400 MB field, the staged copy was 134 ms of a 760 ms call.  Pinning costs about as
 * much as one staged copy, once; the owner's destroy function unpins (bfm_mesh_destroy, bfm_instance_destroy). */
int bfmg_host_pin(void const* ptr, size_t bytes) {
	if (!bfmg_ready() || ptr == nullptr || bytes == 0) {
		return -1;
	}

	if (is_pinned(ptr, bytes)) {
		return 0;
	}

	bfmg_host_unpin(ptr); /* same buffer, grown */

	for (auto& pin : G.pins) {
		if (pin.base == nullptr) {
			if (cudaHostRegister((void*) ptr, bytes, cudaHostRegisterDefault) != cudaSuccess) {
				cudaGetLastError();
				return -1; /* not fatal: the staged path takes it */
			}

			pin.base = (char*) ptr;
			pin.bytes = bytes;
			return 0;
		}
	}

	return -1; /* table full: staged path */
}

void bfmg_host_unpin(void const* ptr) {
	if (ptr == nullptr || !G.ok) {
		return;
	}

	for (auto& pin : G.pins) {
		if (pin.base == (char const*) ptr) {
			cudaSetDevice(G.device);
			cudaStreamSynchronize(G.stream);
			cudaHostUnregister(pin.base);
			cudaGetLastError();
			pin.base = nullptr;
			pin.bytes = 0;
		}
	}
}

int bfmg_upload(void* d_dst, void const* src, size_t bytes) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (bytes == 0) {
		return 0;
	}

	if (is_pinned(src, bytes)) {
		/* page-locked in place: the DMA reads the caller's buffer while the call has already returned - callers
		 * that pin (job.c) do not touch the buffer before the next synchronising call */
		return BFMG_CHECK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, G.stream));
	}

	if (bytes < kStageFrom || !stage_ready()) {
		/* pageable source: the copy is staged by the driver and has returned from the host buffer's point of view
		 * when the call returns */
		return BFMG_CHECK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, G.stream));
	}

	int i = 0;

	for (size_t off = 0; off < bytes; off += kStageBytes, i++) {
		int const b = i & 1;
		size_t const n = bytes - off < kStageBytes ? bytes - off : kStageBytes;

		if (G.stage_busy[b] && BFMG_CHECK(cudaEventSynchronize(G.stage_done[b])) < 0) {
			return -1;
		}

		host_copy(G.stage[b], (char const*) src + off, n);

		if (
			BFMG_CHECK(cudaMemcpyAsync((char*) d_dst + off, G.stage[b], n, cudaMemcpyHostToDevice, G.stream)) < 0 ||
			BFMG_CHECK(cudaEventRecord(G.stage_done[b], G.stream)) < 0
		) {
			return -1;
		}

		G.stage_busy[b] = true;
	}

	return 0; /* the caller's buffer has been read in full; the last chunks are still on their way (stream order) */
}

/* connectivity on its way in: the caller's size_t node numbers narrowed to 32 bits while they are copied into the
 * page-locked staging buffers (one pass over the 1.2 GB of a 50 M-DOF mesh instead of narrow-then-copy-then-DMA).
 * *out_of_range is set when a value is >= limit (the transfer still completes) */
int bfmg_upload_narrow(int32_t* d_dst, size_t const* src, size_t count, size_t limit, int* out_of_range) {
	*out_of_range = 0;

	if (!bfmg_ready()) {
		return -1;
	}

	if (count == 0) {
		return 0;
	}

	bool bad = false;

	if (!stage_ready()) { /* no staging buffers: narrow aside, one plain copy */
		int32_t* const tmp = (int32_t*) malloc(count * sizeof *tmp);

		if (tmp == nullptr) {
			set_error("out of host memory");
			return -1;
		}

		for (size_t i = 0; i < count; i++) {
			bad = bad || src[i] >= limit;
			tmp[i] = (int32_t) src[i];
		}

		int const rv = BFMG_CHECK(cudaMemcpyAsync(d_dst, tmp, count * sizeof *tmp, cudaMemcpyHostToDevice, G.stream)) < 0 || BFMG_CHECK(cudaStreamSynchronize(G.stream)) < 0 ? -1 : 0;

		free(tmp);
		*out_of_range = bad;
		return rv;
	}

	size_t const per_stage = kStageBytes / sizeof(int32_t);
	int i = 0;

	for (size_t off = 0; off < count; off += per_stage, i++) {
		int const b = i & 1;
		size_t const n = count - off < per_stage ? count - off : per_stage;
		int32_t* const out = (int32_t*) G.stage[b];
		size_t const* const in = src + off;

		if (G.stage_busy[b] && BFMG_CHECK(cudaEventSynchronize(G.stage_done[b])) < 0) {
			return -1;
		}

#pragma omp parallel for schedule(static) reduction(|| : bad) if (n >= ((size_t) 1 << 16))
		for (size_t k = 0; k < n; k++) {
			bad = bad || in[k] >= limit;
			out[k] = (int32_t) in[k];
		}

		if (
			BFMG_CHECK(cudaMemcpyAsync(d_dst + off, out, n * sizeof(int32_t), cudaMemcpyHostToDevice, G.stream)) < 0 ||
			BFMG_CHECK(cudaEventRecord(G.stage_done[b], G.stream)) < 0
		) {
			return -1;
		}

		G.stage_busy[b] = true;
	}

	*out_of_range = bad;
	return 0; /* the caller's buffer has been read in full; the last chunks are still on their way (stream order) */
}

int bfmg_download(void* dst, void const* d_src, size_t bytes) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (bytes >= kStageFrom && !is_pinned(dst, bytes) && stage_ready()) {
		size_t const n_chunks = (bytes + kStageBytes - 1) / kStageBytes;

		for (size_t c = 0; c <= n_chunks; c++) {
			if (c < n_chunks) { /* start the DMA of chunk c */
				int const b = (int) (c & 1);
				size_t const off = c * kStageBytes;
				size_t const n = bytes - off < kStageBytes ? bytes - off : kStageBytes;

				if (G.stage_busy[b] && BFMG_CHECK(cudaEventSynchronize(G.stage_done[b])) < 0) {
					return -1;
				}

				if (
					BFMG_CHECK(cudaMemcpyAsync(G.stage[b], (char const*) d_src + off, n, cudaMemcpyDeviceToHost, G.stream)) < 0 ||
					BFMG_CHECK(cudaEventRecord(G.stage_done[b], G.stream)) < 0
				) {
					return -1;
				}

				G.stage_busy[b] = true;
			}

			if (c > 0) { /* ... while the host threads move chunk c - 1 out of its buffer */
				int const b = (int) ((c - 1) & 1);
				size_t const off = (c - 1) * kStageBytes;
				size_t const n = bytes - off < kStageBytes ? bytes - off : kStageBytes;

				if (BFMG_CHECK(cudaEventSynchronize(G.stage_done[b])) < 0) {
					return -1;
				}

				host_copy((char*) dst + off, G.stage[b], n);
				G.stage_busy[b] = false;
			}
		}

		return BFMG_CHECK(cudaStreamSynchronize(G.stream));
	}

	if (bytes != 0 && BFMG_CHECK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, G.stream)) < 0) {
		return -1;
	}

	return BFMG_CHECK(cudaStreamSynchronize(G.stream));
}

int bfmg_copy(void* d_dst, void const* d_src, size_t bytes) {
	if (!bfmg_ready()) {
		return -1;
	}

	return bytes ? BFMG_CHECK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, G.stream)) : 0;
}

int bfmg_zero(void* d_dst, size_t bytes) {
	if (!bfmg_ready()) {
		return -1;
	}

	return bytes ? BFMG_CHECK(cudaMemsetAsync(d_dst, 0, bytes, G.stream)) : 0;
}

int bfmg_sync(void) {
	if (!bfmg_ready()) {
		return -1;
	}

	return BFMG_CHECK(cudaStreamSynchronize(G.stream));
}

int bfmg_tick(void) {
	if (!bfmg_ready()) {
		return -1;
	}

	int const slot = G.next_event;
	G.next_event = (G.next_event + 1) % Context::kEvents;

	return BFMG_CHECK(cudaEventRecord(G.events[slot], G.stream)) < 0 ? -1 : slot;
}

int bfmg_timer_start(int slot) {
	if (!bfmg_ready() || slot < 0 || slot >= 8) {
		return -1;
	}

	return BFMG_CHECK(cudaEventRecord(G.user[slot][0], G.stream));
}

float bfmg_timer_stop(int slot) {
	if (!bfmg_ready() || slot < 0 || slot >= 8) {
		return -1;
	}

	float ms = -1;

	if (cudaEventRecord(G.user[slot][1], G.stream) != cudaSuccess || cudaEventSynchronize(G.user[slot][1]) != cudaSuccess || cudaEventElapsedTime(&ms, G.user[slot][0], G.user[slot][1]) != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}

	return ms;
}

float bfmg_lap(int from, int to) {
	if (from < 0 || to < 0 || !G.ok) {
		return -1;
	}

	float ms = -1;

	if (cudaEventSynchronize(G.events[to]) != cudaSuccess || cudaEventElapsedTime(&ms, G.events[from], G.events[to]) != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}

	return ms;
}

} // extern "C"
