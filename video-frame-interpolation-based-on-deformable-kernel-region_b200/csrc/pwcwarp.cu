// pwcwarp.cu -- PWCDCNet.warp: backward warp of a feature map by a flow with a validity mask.
//
// Behaviour follows PWCNet/PWCNet.py:159-199: vgrid = pixel grid + flow, normalised as
//   nx = 2 * (x + fx) / max(W - 1, 1) - 1          (the align_corners = True style normalisation, :178-179)
// and sampled with torch.nn.functional.grid_sample (bilinear, zero padding), whose published algorithm (PyTorch,
// aten/src/ATen/native/GridSampler.h: grid_sampler_unnormalize, and cuda/GridSampler.cu: grid_sampler_2d_kernel) is
//   align_corners = True :  ix = ((nx + 1) / 2) * (W - 1)          = x + fx: an exact warp
//   align_corners = False:  ix = ((nx + 1) * W - 1) / 2            = (x + fx) * W / (W - 1) - 0.5
//   corners floor(ix), floor(ix) + 1,  weights (ix_se - ix) * (iy_se - iy) ...,  a corner outside the plane contributes nothing.
// The reference's environment pins torch 1.0.1 (environment.yaml:88,104), whose grid_sample has no align_corners argument
// and behaves as align_corners = True -- the exact warp the PWC-Net weights were trained with; `align_corners = 1` is
// therefore the default of the Python front-end.  Run on torch >= 1.3 the same source line silently becomes
// align_corners = False (features shifted by up to half a pixel, first row / column masked at zero flow);
// `align_corners = 0` reproduces that, for callers who need today's behaviour of the unmodified file.  The mask is
// the same sampling of an all-ones tensor, set to 0 below 0.9999 and to 1 otherwise (:193-194); output = sample * mask.
// All coordinate arithmetic is float32 in the reference's operation order; the mask carries no gradient.
// One thread per pixel computes the geometry once and loops over the channels (the reference runs two grid_sample
// launches over separate tensors plus five elementwise kernels); the backward scatters the feature gradient with REDs
// (as grid_sample's backward does) and accumulates the flow gradient in registers.
#include "common.cuh"

namespace vfidkr {
namespace {

constexpr int BX = 32, BY = 8;

struct WarpGeom {
    int x0, y0;                    // north-west corner (floor), may be outside the plane
    float nw, ne, sw, se;          // bilinear weights
    float tx, ty;                  // ix - floor(ix), iy - floor(iy) (for the flow gradient)
    bool in_nw, in_ne, in_sw, in_se;
    float mask;                    // 0 or 1
};

template <bool AC>
__device__ __forceinline__ WarpGeom warp_geom(int w_i, int h_i, float fx, float fy, int W, int H)
{
    WarpGeom g;
    const float vx = __fadd_rn((float)w_i, fx), vy = __fadd_rn((float)h_i, fy);                        // :176
    const float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, vx), (float)max(W - 1, 1)), 1.0f);           // :178
    const float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, vy), (float)max(H - 1, 1)), 1.0f);           // :179
    // grid_sampler_unnormalize
    const float ix = AC ? __fmul_rn(__fdiv_rn(__fadd_rn(nx, 1.0f), 2.0f), (float)(W - 1))
                        : __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(nx, 1.0f), (float)W), 1.0f), 2.0f);
    const float iy = AC ? __fmul_rn(__fdiv_rn(__fadd_rn(ny, 1.0f), 2.0f), (float)(H - 1))
                        : __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(ny, 1.0f), (float)H), 1.0f), 2.0f);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    // the saturating conversion keeps wild coordinates (and NaN -> 0 weights are NaN anyway) out of int overflow
    g.x0 = (int)fmaxf(fminf(fx0, 1e9f), -1e9f);
    g.y0 = (int)fmaxf(fminf(fy0, 1e9f), -1e9f);
    const float x1f = __fadd_rn(fx0, 1.0f), y1f = __fadd_rn(fy0, 1.0f);
    g.nw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(y1f, iy));
    g.ne = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(y1f, iy));
    g.sw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(iy, fy0));
    g.se = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(iy, fy0));
    g.tx = __fsub_rn(ix, fx0);
    g.ty = __fsub_rn(iy, fy0);
    const bool xin0 = g.x0 >= 0 && g.x0 < W, xin1 = g.x0 + 1 >= 0 && g.x0 + 1 < W;
    const bool yin0 = g.y0 >= 0 && g.y0 < H, yin1 = g.y0 + 1 >= 0 && g.y0 + 1 < H;
    g.in_nw = xin0 && yin0; g.in_ne = xin1 && yin0; g.in_sw = xin0 && yin1; g.in_se = xin1 && yin1;
    float m = 0.0f;                // grid_sample of ones, accumulated in its corner order
    if (g.in_nw) m = __fadd_rn(m, g.nw);
    if (g.in_ne) m = __fadd_rn(m, g.ne);
    if (g.in_sw) m = __fadd_rn(m, g.sw);
    if (g.in_se) m = __fadd_rn(m, g.se);
    g.mask = (m < 0.9999f) ? 0.0f : (m > 0.0f ? 1.0f : m);   // :193-194 (NaN stays NaN)
    return g;
}

// Channel loop: 4 gathers per channel, so what pays is loads in flight -- eight channels unrolled at 64 registers (4 blocks
// per SM) against four at 40 registers (6 blocks): 159 -> 89 us at 8 x 32 x 288 x 496; 16 / 32 channels unrolled are slower
// again (193 / 250 us), and so is full occupancy at 32 registers (141 us) -- profiles/r02/pwc_warp_ab_v1.log.
#ifndef VFIDKR_PWC_MINB
#define VFIDKR_PWC_MINB 4
#endif
#ifndef VFIDKR_PWC_UNROLL
#define VFIDKR_PWC_UNROLL 8
#endif
constexpr int PWC_UNROLL = VFIDKR_PWC_UNROLL;
template <bool AC>
__global__ void __launch_bounds__(BX *BY, VFIDKR_PWC_MINB)
pwcwarp_forward_kernel(const float *__restrict__ x, const float *__restrict__ flo, float *__restrict__ out, int C, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flo + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flo + ((size_t)b * 2 + 1) * HW + pix);
    const WarpGeom g = warp_geom<AC>(w_i, h_i, fx, fy, W, H);
    const float *img = x + (size_t)b * C * HW;
    float *o = out + (size_t)b * C * HW + pix;
    const int a = g.y0 * W + g.x0;   // only dereferenced where the in_* flags allow
#pragma unroll PWC_UNROLL
    for (int c = 0; c < C; ++c) {
        const float *pl = img + (size_t)c * HW;
        float acc = 0.0f;            // grid_sampler_2d_kernel's order: nw, ne, sw, se
        if (g.in_nw) acc += __ldg(pl + a) * g.nw;
        if (g.in_ne) acc += __ldg(pl + a + 1) * g.ne;
        if (g.in_sw) acc += __ldg(pl + a + W) * g.sw;
        if (g.in_se) acc += __ldg(pl + a + W + 1) * g.se;
        st_stream(o + (size_t)c * HW, acc * g.mask);   // :199
    }
}

// gx[corner] += gout * mask * weight;  gflo = (W / (W - 1)) * d(sample)/d(ix), (H / (H - 1)) * d(sample)/d(iy)
template <bool AC>
__global__ void __launch_bounds__(BX *BY, 3)
pwcwarp_backward_kernel(const float *__restrict__ x, const float *__restrict__ flo, const float *__restrict__ gout,
                        float *__restrict__ gx, float *__restrict__ gflo, int C, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flo + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flo + ((size_t)b * 2 + 1) * HW + pix);
    const WarpGeom g = warp_geom<AC>(w_i, h_i, fx, fy, W, H);
    float gix = 0.0f, giy = 0.0f;
    if (g.mask != 0.0f) {
        const float *img = x + (size_t)b * C * HW;
        float *gimg = gx + (size_t)b * C * HW;
        const float *go = gout + (size_t)b * C * HW + pix;
        const int a = g.y0 * W + g.x0;
        constexpr int CH = 4;
        for (int c0 = 0; c0 < C; c0 += CH) {
            float gv[CH], v[CH][4];
#pragma unroll
            for (int k = 0; k < CH; ++k) {   // every load of the chunk before its first RED
                const bool ok = c0 + k < C;
                const float *pl = img + (size_t)(ok ? c0 + k : c0) * HW;
                gv[k] = ok ? ld_stream(go + (size_t)(c0 + k) * HW) * g.mask : 0.0f;
                v[k][0] = g.in_nw ? __ldg(pl + a) : 0.0f;
                v[k][1] = g.in_ne ? __ldg(pl + a + 1) : 0.0f;
                v[k][2] = g.in_sw ? __ldg(pl + a + W) : 0.0f;
                v[k][3] = g.in_se ? __ldg(pl + a + W + 1) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (c0 + k >= C) break;
                float *gp = gimg + (size_t)(c0 + k) * HW;
                if (g.in_nw) red_add(gp + a, gv[k] * g.nw);
                if (g.in_ne) red_add(gp + a + 1, gv[k] * g.ne);
                if (g.in_sw) red_add(gp + a + W, gv[k] * g.sw);
                if (g.in_se) red_add(gp + a + W + 1, gv[k] * g.se);
                // d/d(ix), d/d(iy) of nw*v0 + ne*v1 + sw*v2 + se*v3 (grid_sampler_2d_backward_kernel)
                gix += gv[k] * ((v[k][1] - v[k][0]) * (1.0f - g.ty) + (v[k][3] - v[k][2]) * g.ty);
                giy += gv[k] * ((v[k][2] - v[k][0]) * (1.0f - g.tx) + (v[k][3] - v[k][1]) * g.tx);
            }
        }
    }
    // d(ix)/d(fx): ix = (x + fx) * (W - 1) / max(W - 1, 1)  or  (x + fx) * W / max(W - 1, 1) - 0.5
    st_stream(gflo + ((size_t)b * 2 + 0) * HW + pix, gix * ((float)(AC ? W - 1 : W) / (float)max(W - 1, 1)));
    st_stream(gflo + ((size_t)b * 2 + 1) * HW + pix, giy * ((float)(AC ? H - 1 : H) / (float)max(H - 1, 1)));
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_pwcwarp_forward(const float *x, const float *flow, float *output, int B, int C, int H, int W,
                                      int align_corners, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535 || !x || !flow || !output) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if (align_corners) pwcwarp_forward_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(x, flow, output, C, H, W);
    else               pwcwarp_forward_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(x, flow, output, C, H, W);
    note_launch();
    return check_launch("pwc warp forward");
}

VFIDKR_API int vfidkr_pwcwarp_backward(const float *x, const float *flow, const float *gradoutput, float *gradx,
                                       float *gradflow, int B, int C, int H, int W, int align_corners, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || B > 65535 || !x || !flow || !gradoutput || !gradx || !gradflow) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 30)) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const int e = set_error(cudaMemsetAsync(gradx, 0, sizeof(float) * (size_t)B * C * H * W, s), "clear gradx");
    if (e) return e;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if (align_corners) pwcwarp_backward_kernel<true><<<grid, block, 0, s>>>(x, flow, gradoutput, gradx, gradflow, C, H, W);
    else               pwcwarp_backward_kernel<false><<<grid, block, 0, s>>>(x, flow, gradoutput, gradx, gradflow, C, H, W);
    note_launch();
    return check_launch("pwc warp backward");
}
