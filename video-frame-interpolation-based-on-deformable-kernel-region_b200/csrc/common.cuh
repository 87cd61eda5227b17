// common.cuh -- shared helpers for the sm_100a kernels of libvfidkr_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vfidkr_b200.h"

#define VFIDKR_API extern "C" __attribute__((visibility("default")))

namespace vfidkr {

// ---- host side -----------------------------------------------------------------------
void note_launch(int n = 1);              // relaxed launch counter (capi.cu)
int  check_launch(const char *what);      // cudaGetLastError -> VFIDKR_OK / VFIDKR_ERR_CUDA
int  set_error(cudaError_t e, const char *what);
int  sm_count();                          // cached multiprocessor count of the current device
// stream-ordered scratch memory from a library-private pool (capi.cu): no synchronisation, cached by the pool up to
// VFIDKR_SCRATCH_RETAIN_MB; release with cudaFreeAsync on the same stream.  The device's default pool is left alone.
int  stream_scratch_alloc(void **p, size_t bytes, cudaStream_t s);
// environment switch read once per process (test hooks must not cost a getenv per launch)
const char *env_once(const char *name);

static inline unsigned ceil_div(long long a, long long b) { return (unsigned)((a + b - 1) / b); }
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Division of a non-negative 31-bit integer by a run-time constant without the ~40-instruction
// integer-division sequence: quotient = umulhi(n, mul) >> shr (mul/shr precomputed on the host).
struct FastDiv {
    unsigned mul, shr, div;
    FastDiv() : mul(0), shr(0), div(1) {}
    explicit FastDiv(unsigned d) : div(d)
    {
        if (d <= 1) { mul = 0; shr = 0; div = 1; return; }
        unsigned lg = 31 - __builtin_clz(d);
        if (d & (d - 1)) ++lg;   // ceil(log2(d))
        const unsigned p = 31 + lg;
        mul = (unsigned)(((1ull << p) + d - 1) / d);
        shr = p - 32;
    }
#ifdef __CUDACC__
    __device__ __forceinline__ int quot(int n) const { return div == 1 ? n : (int)(__umulhi((unsigned)n, mul) >> shr); }
#endif
};

// ---- device side ---------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// streaming (read-once) loads / stores: keep them out of L1 so the gather working set stays resident
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// fire-and-forget float add (RED.E.ADD.F32)
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }


// ---- debug build -DVFIDKR_BOUNDS_CHECK (compute-sanitizer is not available on every pool): the kernels that read images
// through shared-memory windows verify, tap by tap, (1) that the shared-memory index lies inside the window ring / region
// and (2) that the value found there is bit-identical to the image value in global memory it stands for -- a stale or
// not-yet-filled window row (a pipeline race) or a wrong index cannot pass.  Counters are per translation unit;
// vfidkr_debug_bounds_counts (capi.cu) adds them up.  Production builds compile none of this.
#ifdef VFIDKR_BOUNDS_CHECK
namespace {
__device__ unsigned long long g_bounds_counts[2];   // [0] checks executed, [1] checks failed (this translation unit)
}
__device__ __forceinline__ void bounds_check(bool ok)
{
    const unsigned m = __activemask(), bad = __ballot_sync(m, !ok);
    unsigned lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    if (lane == (unsigned)(__ffs(m) - 1)) {
        atomicAdd(&g_bounds_counts[0], (unsigned long long)__popc(m));
        if (bad) atomicAdd(&g_bounds_counts[1], (unsigned long long)__popc(bad));
    }
}
#define VFIDKR_BOUNDS_ACCESSOR(name)                                                                             \
    namespace vfidkr { int name(unsigned long long *out2) {                                                      \
        return cudaMemcpyFromSymbol(out2, g_bounds_counts, 2 * sizeof(unsigned long long)) != cudaSuccess; } }
#else
#define VFIDKR_BOUNDS_ACCESSOR(name)
#endif

}  // namespace vfidkr
