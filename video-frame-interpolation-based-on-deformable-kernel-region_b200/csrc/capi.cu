// capi.cu -- library info entry points and shared host helpers
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace vfidkr {

static std::atomic<unsigned long long> g_launches{0};
static thread_local char t_error[256] = "";

void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int set_error(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return VFIDKR_OK;
    snprintf(t_error, sizeof t_error, "%s: %s", what, cudaGetErrorString(e));
    return VFIDKR_ERR_CUDA;
}

int check_launch(const char *what) { return set_error(cudaGetLastError(), what); }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn resolve_encode()
{
    static std::atomic<void *> cached{nullptr};
    void *fn = cached.load(std::memory_order_acquire);
    if (fn) return (encode_tiled_fn)fn;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) {
        (void)cudaGetLastError();
        return nullptr;
    }
    cached.store(fn, std::memory_order_release);
    return (encode_tiled_fn)fn;
}

// Encoding a tensor map is a driver call; callers hit the same (base, shape, box) repeatedly (the allocator
// recycles blocks), so a small thread-local direct-mapped cache removes it from the steady-state launch path.
struct MapKey {
    const float *base; uint64_t W, H, D; uint32_t bw, bh, bd;
    bool operator==(const MapKey &o) const
    { return base == o.base && W == o.W && H == o.H && D == o.D && bw == o.bw && bh == o.bh && bd == o.bd; }
};
struct MapSlot { MapKey key; CUtensorMap map; bool valid; };

bool encode_tensor_map_3d(CUtensorMap *map, const float *base, uint64_t W, uint64_t H, uint64_t D,
                          uint32_t boxW, uint32_t boxH, uint32_t boxD)
{
    constexpr int SLOTS = 64;
    static thread_local MapSlot cache[SLOTS];
    const MapKey key{base, W, H, D, boxW, boxH, boxD};
    const uint64_t h = (reinterpret_cast<uintptr_t>(base) >> 8) * 0x9E3779B97F4A7C15ull ^ (W * 31 + H * 17 + D * 7 + boxD);
    MapSlot &slot = cache[(h >> 32) % SLOTS];
    if (slot.valid && slot.key == key) { *map = slot.map; return true; }

    encode_tiled_fn enc = resolve_encode();
    if (!enc) return false;
    const cuuint64_t dims[3] = {W, H, D};
    const cuuint64_t strides[2] = {W * sizeof(float), W * H * sizeof(float)};   // byte strides of dims 1 and 2
    const cuuint32_t box[3] = {boxW, boxH, boxD};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    slot.key = key; slot.map = *map; slot.valid = true;
    return true;
}

}  // namespace vfidkr

VFIDKR_API int vfidkr_abi_version(void) { return 100; }
VFIDKR_API unsigned long long vfidkr_launch_count(void)
{
    return vfidkr::g_launches.load(std::memory_order_relaxed);
}
VFIDKR_API const char *vfidkr_last_error(void) { return vfidkr::t_error; }
