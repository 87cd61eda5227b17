// interpolation.cu -- Interpolation / InterpolationCh: bilinear backward warp with zero fill.
//
// Behaviour follows my_package/Interpolation/interpolation_cuda_kernel.cu:29-204 (InterpolationCh is
// textually the same kernel).  The four corner offsets and blend weights are computed once per pixel
// and reused for every channel; the backward makes one pass over the channels (the reference makes
// three) and stores the flow gradient once.  gradinput1 really scatters, so it keeps RED atomics;
// clamped corners that coincide are merged into one RED.
#include "common.cuh"

namespace vfidkr {
namespace {

constexpr int BX = 32, BY = 8;

struct Bilin {
    bool in_range;
    int aTL, aTR, aBL, aBR;
    float alpha, beta, gam1, gam2;
};

__device__ __forceinline__ Bilin bilin(int w_i, int h_i, float fx, float fy, int W, int H)
{
    Bilin r;
    const float x2 = __fadd_rn((float)w_i, fx), y2 = __fadd_rn((float)h_i, fy);
    r.in_range = x2 >= 0.0f && y2 >= 0.0f && x2 < (float)W && y2 < (float)H;   // strict upper bound (:71)
    const int L = (int)x2, T = (int)y2;
    const int R = min(L + 1, W - 1), Bm = min(T + 1, H - 1);                    // :73-76
    r.alpha = __fsub_rn(x2, (float)L);                                          // :77
    r.beta = __fsub_rn(y2, (float)T);                                           // :78
    r.gam1 = __fsub_rn((float)Bm, y2);   // backward: gamma = iy2_B - y2 with the CLAMPED corner (:163)
    r.gam2 = __fsub_rn((float)R, x2);    // :180
    r.aTL = T * W + L; r.aTR = T * W + R; r.aBL = Bm * W + L; r.aBR = Bm * W + R;
    return r;
}

// CT > 0: compile-time channel count -- the 4 * CT gathers of a pixel are independent and all in flight together;
// CT == 0: run-time C, one channel at a time (keeps the register count low enough for full occupancy either way).
template <int CT>
__global__ void __launch_bounds__(BX *BY, 8)
interp_forward_kernel(const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                      int Crt, int H, int W)
{
    const int C = CT > 0 ? CT : Crt;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
    const Bilin g = bilin(w_i, h_i, fx, fy, W, H);
    const float *img = in1 + (size_t)b * C * HW;
    float *o = out + (size_t)b * C * HW + pix;
    if (!g.in_range) {
        for (int c = 0; c < C; ++c) st_stream(o + (size_t)c * HW, 0.0f);   // :88-92
        return;
    }
    const float wTL = (1 - g.alpha) * (1 - g.beta), wTR = g.alpha * (1 - g.beta);
    const float wBL = (1 - g.alpha) * g.beta, wBR = g.alpha * g.beta;
    if (CT > 0) {
        float r[CT > 0 ? CT : 1];
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float *pl = img + (size_t)c * HW;
            r[c] = wTL * __ldg(pl + g.aTL) + wTR * __ldg(pl + g.aTR) + wBL * __ldg(pl + g.aBL) + wBR * __ldg(pl + g.aBR);   // :85-86
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) st_stream(o + (size_t)c * HW, r[c]);
    } else {
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const float *pl = img + (size_t)c * HW;
            st_stream(o + (size_t)c * HW,
                      wTL * __ldg(pl + g.aTL) + wTR * __ldg(pl + g.aTR) + wBL * __ldg(pl + g.aBL) + wBR * __ldg(pl + g.aBR));
        }
    }
}

// CT > 0: compile-time channel count -- every load of the pixel (upstream gradient + 4 corners per channel) is
// requested before the first RED (a load issued after a RED waits behind it), then the REDs go out back to back.
// CT == 0: run-time C in chunks of 4 with the same order.  The register cap keeps four blocks resident (six would spill): the first
// version (79 registers, 3 blocks, four-way branchy merge of clamped corners, 390 instructions per warp) was
// latency-bound at 585 us for 1080p x 8.  (The run-time-C variant keeps 96 bytes of spills under this cap; two resident
// blocks without spills were measured slower: 562 vs 446 us at 4 x 5 x 1152 x 1984.)
template <int CT>
__global__ void __launch_bounds__(BX *BY, 4)
interp_backward_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ gout,
                       float *__restrict__ gi1, float *__restrict__ gi2, int Crt, int H, int W)
{
    const int C = CT > 0 ? CT : Crt;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
    const Bilin g = bilin(w_i, h_i, fx, fy, W, H);
    float bx = 0.0f, by = 0.0f;
    if (g.in_range) {
        const float *img = in1 + (size_t)b * C * HW;
        float *gimg = gi1 + (size_t)b * C * HW;
        const float *go = gout + (size_t)b * C * HW + pix;
        const float wTL = (1 - g.alpha) * (1 - g.beta), wTR = g.alpha * (1 - g.beta);
        const float wBL = (1 - g.alpha) * g.beta, wBR = g.alpha * g.beta;
        constexpr int CH = CT > 0 ? CT : 4;
        for (int c0 = 0; c0 < C; c0 += CH) {
            float gv[CH], TL[CH], TR[CH], BL[CH], BR[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const bool ok = c0 + k < C;
                const float *pl = img + (size_t)(ok ? c0 + k : c0) * HW;
                gv[k] = ok ? ld_stream(go + (size_t)(c0 + k) * HW) : 0.0f;
                TL[k] = __ldg(pl + g.aTL); TR[k] = __ldg(pl + g.aTR); BL[k] = __ldg(pl + g.aBL); BR[k] = __ldg(pl + g.aBR);
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (c0 + k >= C) break;
                float *gp = gimg + (size_t)(c0 + k) * HW;
                // :156-159 -- four atomics; corners that coincide through the border clamp simply hit the same cell twice
                red_add(gp + g.aTL, gv[k] * wTL);
                red_add(gp + g.aTR, gv[k] * wTR);
                red_add(gp + g.aBL, gv[k] * wBL);
                red_add(gp + g.aBR, gv[k] * wBR);
                bx += gv[k] * (g.gam1 * (TR[k] - TL[k]) + (1 - g.gam1) * (BR[k] - BL[k]));   // :165-173
                by += gv[k] * (g.gam2 * (BL[k] - TL[k]) + (1 - g.gam2) * (BR[k] - TR[k]));   // :182-190
            }
        }
    }
    st_stream(gi2 + ((size_t)b * 2 + 0) * HW + pix, bx);
    st_stream(gi2 + ((size_t)b * 2 + 1) * HW + pix, by);
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_interpolation_forward(const float *input1, const float *input2, float *output,
                                            int B, int C, int H, int W, int require_c3, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535 || !input1 || !input2 || !output) return VFIDKR_ERR_ARG;
    if (require_c3 && C != 3) return VFIDKR_ERR_ARG;   // interpolation_cuda.cc:19
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if (C == 3) interp_forward_kernel<3><<<grid, block, 0, (cudaStream_t)stream>>>(input1, input2, output, C, H, W);
    else        interp_forward_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(input1, input2, output, C, H, W);
    note_launch();
    return check_launch("interpolation forward");
}

VFIDKR_API int vfidkr_interpolation_backward(const float *input1, const float *input2, const float *gradoutput,
                                             float *gradinput1, float *gradinput2,
                                             int B, int C, int H, int W, int require_c3, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !gradoutput || !gradinput1 || !gradinput2) return VFIDKR_ERR_ARG;
    if (require_c3 && C != 3) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    int e = set_error(cudaMemsetAsync(gradinput1, 0, sizeof(float) * (size_t)B * C * H * W, s), "clear gradinput1");
    if (e) return e;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if (C == 3) interp_backward_kernel<3><<<grid, block, 0, s>>>(input1, input2, gradoutput, gradinput1, gradinput2, C, H, W);
    else        interp_backward_kernel<0><<<grid, block, 0, s>>>(input1, input2, gradoutput, gradinput1, gradinput2, C, H, W);
    note_launch();
    return check_launch("interpolation backward");
}
