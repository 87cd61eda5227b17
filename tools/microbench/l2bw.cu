// l2bw.cu -- microbenchmark: L2 -> SM read bandwidth for an L2-resident buffer (ld.global.cg, 128-bit) and the
// HBM read bandwidth for a buffer far larger than L2.  Sizing input for the shared-memory staging design of the
// FilterInterpolation kernels: how much halo re-read traffic the L2 can absorb.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/l2bw l2bw.cu && ./_build/l2bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) read_kernel(const float4 *__restrict__ p, size_t n4, int reps, float *sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n4; i += 8 * stride) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcg(p + i + k * stride);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
        }
        for (; i < n4; i += stride) { float4 v = __ldcg(p + i); acc += v.x + v.y + v.z + v.w; }
    }
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    float *buf, *sink;
    const size_t big = (size_t)2 << 30;
    cudaMalloc(&buf, big);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, big);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t sizes_mb[] = {8, 16, 32, 64, 96, 256, 2048};
    for (size_t mb : sizes_mb) {
        const size_t n4 = mb * 1024 * 1024 / 16;
        const int reps = mb <= 96 ? 40 : (mb == 256 ? 10 : 3);
        for (int ctas_per_sm : {2, 4, 8}) {
            read_kernel<<<148 * ctas_per_sm, 256>>>((const float4 *)buf, n4, 2, sink);   // warm
            cudaEventRecord(a);
            read_kernel<<<148 * ctas_per_sm, 256>>>((const float4 *)buf, n4, reps, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            printf("buffer %5zu MB  %d CTAs/SM  %8.1f GB/s\n", mb, ctas_per_sm, (double)mb * 1048576.0 * reps / (ms * 1e-3) / 1e9);
        }
    }
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
