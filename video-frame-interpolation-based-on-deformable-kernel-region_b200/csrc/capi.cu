// capi.cu -- library info entry points and shared host helpers
#include <atomic>
#include <cstdio>
#include <cstring>

#include "common.cuh"

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

}  // namespace vfidkr

VFIDKR_API int vfidkr_abi_version(void) { return 100; }
VFIDKR_API unsigned long long vfidkr_launch_count(void)
{
    return vfidkr::g_launches.load(std::memory_order_relaxed);
}
VFIDKR_API const char *vfidkr_last_error(void) { return vfidkr::t_error; }
