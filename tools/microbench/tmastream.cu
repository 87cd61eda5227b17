// tmastream.cu -- microbenchmark: HBM read bandwidth of TMA box loads in the access pattern of the FilterInterpolation
// strip kernel (one persistent CTA per SM walks 128-column strips downwards; per tile one 3-D box [BW cols x BH rows x
// BP planes] of a [W x H x P] fp32 tensor goes into a STAGES-deep shared-memory ring), with no compute at all.  It bounds
// what fi_strip.cu can reach for its filter stream and shows how the box shape changes that bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../video-frame-*/csrc -I ../../include -o _build/tmastream \
//        tmastream.cu ../../video-frame-*/csrc/capi.cu -lcuda   (see run line in DESIGN.md)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "tma.cuh"

using namespace vfidkr;

template <int BW, int BH, int BP, int STAGES>
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int tiles_x, int tiles_y, int planes_groups, int nitems, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    constexpr int STAGE_FLOATS = BW * BH * BP;
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + STAGES * STAGE_FLOATS);
    uint64_t *empty = full + STAGES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
        fence_mbar_init();
    }
    __syncthreads();
    // item = (plane group g, column block bx): the CTA walks all tiles_y tiles of the strip downwards
    const int my_items = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n = my_items * tiles_y;
    float acc = 0.f;
    if (tid < 32) {   // producer warp
        if (tid == 0)
            for (int t = 0; t < n; ++t) {
                const int item = blockIdx.x + (t / tiles_y) * gridDim.x, ty = t % tiles_y;
                const int bx = item % tiles_x, g = item / tiles_x;
                const int s = t % STAGES;
                if (t >= STAGES) mbar_wait(&empty[s], (uint32_t)(((t / STAGES) - 1) & 1));
                mbar_arrive_expect_tx(&full[s], STAGE_FLOATS * 4);
                tma_load_3d(ring + s * STAGE_FLOATS, &map, &full[s], bx * BW, ty * BH, g * BP);
            }
    } else {          // three consumer warps: wait, touch one value, release
        for (int t = 0; t < n; ++t) {
            const int s = t % STAGES;
            mbar_wait(&full[s], (uint32_t)((t / STAGES) & 1));
            acc += ring[s * STAGE_FLOATS + tid];
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int BW, int BH, int BP, int STAGES>
void run(const float *buf, int W, int H, int P, float *sink)
{
    CUtensorMap map;
    if (!encode_tensor_map_3d(&map, buf, W, H, P, BW, BH, BP)) { printf("encode failed\n"); return; }
    const int tiles_x = (W + BW - 1) / BW, tiles_y = (H + BH - 1) / BH, groups = P / BP;
    const int nitems = tiles_x * groups;
    const size_t smem = (size_t)STAGES * BW * BH * BP * 4 + 256;
    auto k = stream_kernel<BW, BH, BP, STAGES>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148, 128, smem>>>(map, tiles_x, tiles_y, groups, nitems, sink);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) k<<<148, 128, smem>>>(map, tiles_x, tiles_y, groups, nitems, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)W * H * P * 4;
    printf("box %3d x %d x %2d, %d stages (%3zu KB in flight/SM): %7.1f us  %6.0f GB/s  (%s)\n", BW, BH, BP, STAGES,
           smem / 1024, ms / 5 * 1e3, bytes / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const int W = 1984, H = 1152, P = 128;   // the filter tensor of the bench: 8 frames x 16 planes, 1.17 GB
    float *buf, *sink;
    cudaMalloc(&buf, (size_t)W * H * P * 4);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, (size_t)W * H * P * 4);
    run<128, 4, 16, 4>(buf, W, H, P, sink);   // fi_strip.cu's filter stream
    run<128, 4, 16, 6>(buf, W, H, P, sink);
    run<128, 8, 16, 3>(buf, W, H, P, sink);
    run<128, 8, 8, 6>(buf, W, H, P, sink);
    run<256, 4, 8, 6>(buf, W, H, P, sink);
    run<256, 2, 16, 6>(buf, W, H, P, sink);
    run<128, 16, 4, 6>(buf, W, H, P, sink);
    run<64, 8, 16, 6>(buf, W, H, P, sink);
    return 0;
}
