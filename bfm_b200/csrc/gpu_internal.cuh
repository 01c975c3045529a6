/*
 * gpu_internal.cuh - helpers shared by the .cu translation units (not part of any ABI).
 */
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "gpu.h"

#define BFMG_HIDDEN __attribute__((visibility("hidden")))

BFMG_HIDDEN bool bfmg_ready();                 /* context initialised and current */
BFMG_HIDDEN cudaStream_t bfmg_stream();
BFMG_HIDDEN void bfmg_count_launch(size_t n);
BFMG_HIDDEN void* bfmg_pinned();               /* 4 KiB page-locked scratch */
BFMG_HIDDEN cudaEvent_t bfmg_poll_event(int i); /* two untimed events */
BFMG_HIDDEN int bfmg_check(cudaError_t rc, char const* what, char const* file, int line);
BFMG_HIDDEN void bfmg_set_error(char const* fmt, ...);
BFMG_HIDDEN int bfmg_device();

/* D^-1/2 of the diagonal, b^ = D^-1/2 b, A^ = D^-1/2 A D^-1/2 into d_scaled (solver.cu) */
BFMG_HIDDEN int bfmg_scale_system(bfmg_pattern_t const* pat, double const* d_val, double const* d_b, double* d_dscale, double* d_bhat, double* d_scaled);

/* NVLink peer-memory mailboxes (p2p.cuh, dist.cu): NULL when unavailable */
struct P2p;
BFMG_HIDDEN P2p const* bfmg_dist_p2p();
BFMG_HIDDEN int bfmg_dist_p2p_begin(int fits, int* all_fit); /* collective: zero the round words, agree on using peer memory */
BFMG_HIDDEN int bfmg_dist_p2p_failed();

#define BFMG_CHECK(call) bfmg_check((call), #call, __FILE__, __LINE__)

/* Programmatic dependent launch (BFM_PDL=1; off by default).  A multigrid-preconditioned iteration is a string of
 * ~200 kernels, most of them a few microseconds long, so what they cost is the gap between the end of one and the
 * start of the next.  Every kernel of this library begins with pdl_sync() - it lets the NEXT kernel of the stream be
 * scheduled at once (griddepcontrol.launch_dependents, PREEXIT in SASS) and then waits until the PREVIOUS one has
 * finished and its writes are visible (griddepcontrol.wait, ACQBULK) - and with BFM_PDL=1 it is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so that its launch latency overlaps the predecessor's execution;
 * stream capture keeps the edges.  Nothing may be read before pdl_sync(): it is the first statement of every kernel,
 * and a no-op in a classic launch.  Measured on the B200 (profiles/r2_summary.md): 915 vs 914 ms per 50 M-DOF step,
 * 57.4 vs 57.1 ms at 2 M DOF - inside a replayed CUDA graph the launches are already back to back, so the default
 * stays the classic launch. */
__device__ __forceinline__ void pdl_sync() {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

BFMG_HIDDEN bool bfmg_pdl(); /* context.cu: programmatic launches on (default) */

template <typename... Params, typename... Args>
static inline cudaError_t bfmg_launch(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
	cudaLaunchConfig_t config = {};
	cudaLaunchAttribute attribute[1];

	attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attribute[0].val.programmaticStreamSerializationAllowed = 1;

	config.gridDim = grid;
	config.blockDim = block;
	config.dynamicSmemBytes = smem;
	config.stream = bfmg_stream();
	config.attrs = attribute;
	config.numAttrs = bfmg_pdl() ? 1 : 0;

	return cudaLaunchKernelEx(&config, kernel, static_cast<Params>(args)...);
}

/* launch on the library stream, count it, report configuration errors */
#define BFMG_LAUNCH(kernel, grid, block, smem, ...)                       \
	(bfmg_count_launch(1), bfmg_check(bfmg_launch((kernel), dim3(grid), dim3(block), (smem), __VA_ARGS__), #kernel, __FILE__, __LINE__))

constexpr int kWarp = 32;
constexpr int kBlock = 256;                  /* 8 warps per CTA for every kernel in this library */
constexpr int kWarpsPerBlock = kBlock / kWarp;

/* persistent-style grids: a multiple of the SM count, capped by the amount of work */
static inline int bfmg_grid(int64_t work_items_per_block_units, int ctas_per_sm) {
	int64_t const cap = (int64_t) bfmg_sm_count() * ctas_per_sm;
	int64_t g = work_items_per_block_units < cap ? work_items_per_block_units : cap;
	return (int) (g < 1 ? 1 : g);
}

/* resident CTAs of `kernel` per SM at kBlock threads: grids are sized to exactly one full wave
 * (SMs x resident CTAs) so that grid-stride loops split the work evenly - 8 CTAs/SM of a kernel that
 * fits only 6 would run as 1.33 waves with a 1/3-occupied tail (seen in the first ncu capture) */
template <typename Kernel>
static inline int bfmg_resident_ctas(Kernel kernel) {
	int n = 0;

	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBlock, 0) != cudaSuccess || n < 1) {
		cudaGetLastError();
		n = 1;
	}

	return n;
}

/* streaming 16-byte loads that do not pollute L1 (matrix values and column indices are read once per
 * SpMV; L1 is kept for the gathered vector) */
__device__ __forceinline__ double2 ld_stream(double2 const* p) {
	double2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
	return v;
}

__device__ __forceinline__ float2 ld_stream(float2 const* p) {
	float2 v;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
	return v;
}

__device__ __forceinline__ double ld_stream(double const* p) {
	double v;
	asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ float ld_stream(float const* p) {
	float v;
	asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ int ld_stream(int const* p) {
	int v;
	asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int off = kWarp / 2; off > 0; off >>= 1) {
		v += __shfl_down_sync(0xffffffffu, v, off);
	}

	return v;
}
