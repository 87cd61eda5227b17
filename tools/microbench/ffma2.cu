// ffma2.cu -- microbenchmark: fp32 FMA issue rate of sm_100a for (a) the three-register FFMA, (b) the packed FFMA2
// (PTX fma.rn.f32x2, two FMAs per lane and instruction), in the operand pattern of the correlation kernel: an
// accumulator tile in registers, one operand broadcast per row of the tile.  Also (c) vector RED (red.global.add.v4.f32)
// throughput into an L2-resident image, identity mapping (one RED per thread and cell), the projection splat's pattern.
// Sizing input for the correlation / SeparableConv register tiles and the projection splat.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _build/ffma2 ffma2.cu && ./_build/ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

constexpr int ACC = 64;   // accumulators per thread

__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float seed)
{
    float acc[ACC], a[8], b[8];
    for (int i = 0; i < ACC; ++i) acc[i] = (float)i;
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; b[i] = seed * 0.5f - i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-7f; }
    }
    float s = 0.f;
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float *out, int iters, float seed)
{
    unsigned long long acc[ACC / 2], a[8], b[4];
    for (int i = 0; i < ACC / 2; ++i) acc[i] = pack((float)i, (float)i + 0.5f);
    for (int i = 0; i < 8; ++i) a[i] = pack(seed + i + threadIdx.x, seed + i + threadIdx.x);   // broadcast pair
    for (int i = 0; i < 4; ++i) b[i] = pack(seed * 0.5f - 2 * i, seed * 0.5f - 2 * i - 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(acc[i * 4 + j], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) ffma2(b[i], b[i], pack(1.0000001f, 1.0000001f));   // keeps the loop from being hoisted
    }
    float s = 0.f;
    for (int i = 0; i < ACC / 2; ++i) { float2 v = *reinterpret_cast<float2 *>(&acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) red4_kernel(float4 *img, size_t cells, int reps, int shift)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
            size_t j = i + (size_t)shift * (r + 1);
            if (j >= cells) j -= cells;
            atomicAdd(img + j, make_float4(1.f, 2.f, 3.f, 0.f));
        }
}
__global__ void __launch_bounds__(256) red1_kernel(float *img, size_t cells, int reps)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) atomicAdd(img + i, 1.f);
}

template <typename F>
static float time_ms(F f, int n = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < n; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / n;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        const float t1 = time_ms([&] { ffma_kernel<<<sms * ctas, 256>>>(out, iters, 1.5f); });
        const float t2 = time_ms([&] { ffma2_kernel<<<sms * ctas, 256>>>(out, iters, 1.5f); });
        const double fma1 = (double)sms * ctas * 256 * iters * 64, fma2 = (double)sms * ctas * 256 * iters * (64 + 8);
        printf("%d CTAs/SM x 256 thr: FFMA %.1f TFLOP/s (%.1f FMA/clk/SM at %d MHz nominal), FFMA2 %.1f TFLOP/s (%.1f FMA/clk/SM)\n", ctas,
               2 * fma1 / t1 / 1e9, fma1 / (t1 * 1e-3) / sms / (khz * 1e3), khz / 1000, 2 * fma2 / t2 / 1e9, fma2 / (t2 * 1e-3) / sms / (khz * 1e3));
    }
    // vector RED into an L2-resident image (36.5 MB = one 1080p frame of 16 B cells) and into a 292 MB one
    for (size_t cells : {(size_t)1152 * 1984, (size_t)8 * 1152 * 1984}) {
        float4 *img;
        cudaMalloc(&img, cells * sizeof(float4));
        cudaMemset(img, 0, cells * sizeof(float4));
        const int reps = cells > 4000000 ? 2 : 16;
        for (int ctas : {2, 4, 8}) {
            for (int shift : {0, 7777}) {
                const float t = time_ms([&] { red4_kernel<<<sms * ctas, 256>>>(img, cells, reps, shift); });
                printf("RED.v4 %zu MB image, %d CTAs/SM, shift %d: %.1f G RED/s (%.1f us per 2.29 M REDs)\n", cells * 16 >> 20, ctas, shift,
                       (double)cells * reps / t / 1e6, t * 1e3 * 2285568.0 / ((double)cells * reps));
            }
        }
        const float t1 = time_ms([&] { red1_kernel<<<sms * 8, 256>>>(reinterpret_cast<float *>(img), cells * 4, reps); });
        printf("RED.f32 scalar, %zu MB image: %.1f G RED/s\n", cells * 16 >> 20, (double)cells * 4 * reps / t1 / 1e6);
        cudaFree(img);
    }
    return 0;
}
