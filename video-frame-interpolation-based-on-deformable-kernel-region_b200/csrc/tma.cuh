// tma.cuh -- Tensor Memory Accelerator (cp.async.bulk.tensor) and mbarrier helpers, sm_100a.
//
// Host side: tensor maps are encoded with cuTensorMapEncodeTiled, resolved through
// cudaGetDriverEntryPoint so the library does not link libcuda.
// Device side: thin inline-PTX wrappers (the SASS shows UTMALDG / SYNCS for these).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vfidkr {

// Encodes a rank-3 fp32 tiled tensor map over a contiguous [D][H][W] array (W fastest).
// Requirements (checked by the caller): base 16-byte aligned, W % 4 == 0, box dims <= 256,
// boxW * 4 a multiple of 16.  Returns false if the driver entry point is unavailable or rejects the map.
bool encode_tensor_map_3d(CUtensorMap *map, const float *base, uint64_t W, uint64_t H, uint64_t D,
                          uint32_t boxW, uint32_t boxH, uint32_t boxD);
// Rank-4 variant over a contiguous [N][C][H][W] array: a box that runs past C (or H, W) is zero-filled
// instead of running into the next batch item.
bool encode_tensor_map_4d(CUtensorMap *map, const float *base, uint64_t W, uint64_t H, uint64_t C, uint64_t N,
                          uint32_t boxW, uint32_t boxH, uint32_t boxC);
// same with padded rows: extent W (reads past it are zero-filled), row pitch `row_pitch` elements (a multiple of 4)
bool encode_tensor_map_4d_pitched(CUtensorMap *map, const float *base, uint64_t W, uint64_t H, uint64_t C, uint64_t N,
                                  uint64_t row_pitch, uint32_t boxW, uint32_t boxH, uint32_t boxC);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// plain arrival (consumer -> producer "slot free" signalling)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait with a watchdog: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
// try_wait suspends the thread for a hardware-defined time slice per poll, so the bound is far above any
// legitimate wait (seconds) yet finite.
__device__ __forceinline__ void mbar_wait_guarded(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}

// try_wait with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the hint
// expires, so a polling loop costs a handful of instructions per microsecond instead of per ~20 ns.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// non-blocking test of a phase
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// guarded wait built on the hinted try_wait (see mbar_wait_guarded)
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait_hint(bar, parity, 20000u); ++spins)
        if (spins > (1u << 22)) __trap();
}

// ---- the same primitives on precomputed 32-bit shared addresses (hot loops: no address conversion per call) ----
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait of a compute warp: a failed probe is followed by a real sleep, so a warp that is ahead of its data does not
// burn the issue slots of the warps it is waiting for.  Traps instead of hanging on a protocol error.
__device__ __forceinline__ void mbar_wait_backoff_a(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait_a(bar, parity)) return;
    for (uint32_t spins = 0;; ++spins) {
        __nanosleep(128);
        if (mbar_try_wait_a(bar, parity)) return;
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// shared-memory min / max without the compiler's warp-aggregation wrapper (the caller is a single elected lane)
__device__ __forceinline__ void red_shared_min_a(uint32_t addr, int v)
{
    asm volatile("red.shared::cta.min.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_max_a(uint32_t addr, int v)
{
    asm volatile("red.shared::cta.max.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// 3-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
// L2 prefetch of a 3-D box: brings the bytes from HBM into L2 without occupying shared memory, so a later
// tma_load_3d of the same box pays L2 latency only (decouples HBM latency from the depth of the smem ring)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
// 4-D tiled load (x, y, channel, batch)
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int c, int n)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n)
        : "memory");
}
// same with an L2 cache-policy hint (createpolicy-generated 64-bit policy)
__device__ __forceinline__ void tma_load_3d_hint(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                                 int z, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Ampere-style asynchronous 4-byte copy global -> shared with zero fill (src_bytes = 0 writes 0.0f and
// reads nothing); used where TMA's 16-byte pitch requirement does not hold.  SASS: LDGSTS.
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gmem_src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 4u : 0u)
                 : "memory");
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, bool valid)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 16u : 0u)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace vfidkr
