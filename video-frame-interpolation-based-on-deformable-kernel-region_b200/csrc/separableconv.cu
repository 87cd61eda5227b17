// separableconv.cu -- SeparableConv (per-pixel separable F x F filtering on the valid region) and
// SeparableConvFlow (filter centroid -> flow), forward and backward, for sm_100a.
//
// Behaviour follows my_package/SeparableConv/separableconv_cuda_kernel.cu:29-135 and
// my_package/SeparableConvFlow/separableconvflow_cuda_kernel.cu:29-174.
// Differences in HOW:
//   * F = 51 makes SeparableConv arithmetic-bound (3*C*F*F flops against ~430 bytes per pixel), and what limits a
//     direct kernel is the load per multiply-add.  The tiled kernels stage the image region of a 64 x 16 pixel tile
//     in shared memory, give a thread a 2 x 2 block of pixels and hold the horizontal taps in registers, so that one
//     8-byte shared-memory access feeds eight multiply-adds (forward, filter gradients) or carries eight terms
//     (image gradient).  Measured at 8 x 3 x 256 x 448, F = 51: forward 1.37 -> 0.56 ms, backward 8.2 -> 1.9 ms.
//   * gradinput2 / gradinput3 use no atomics (thread-private sums; the reference issues 2*C*F*F atomicAdds per pixel,
//     :124-127); gradinput1 is summed per tile in shared memory and added to the image with ~9 atomics per element
//     (the reference: F*F per element, :122-123).  The direct kernels (F too large for the shared-memory region) use
//     no atomics at all: gradinput1 is a gather over the output pixels whose window covers the input pixel.
//   * every output element is written (gradinput1 is zeroed by the launcher), so no caller zero-fill is needed.
#include "common.cuh"

namespace vfidkr {
namespace {

constexpr int BX = 32, BY = 8;

// out[b,c,h,w] = sum_y sum_x I[b,c,h+y,w+x] * v[b,y,h,w] * hz[b,x,h,w]      (:65-77)
__global__ void __launch_bounds__(BX *BY)
sepconv_forward_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                       float *__restrict__ out, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= Wo || h_i >= Ho) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo, po = (size_t)h_i * Wo + w_i;
    const float *v = in2 + (size_t)b * F * HWo + po, *hz = in3 + (size_t)b * F * HWo + po;
    for (int c0 = 0; c0 < C; c0 += 3) {   // three channels per pass share the v/hz loads
        float acc[3] = {0.f, 0.f, 0.f};
        const float *pl = in1 + ((size_t)b * C + c0) * HW + (size_t)h_i * W + w_i;
        for (int y = 0; y < F; ++y) {
            const float vy = __ldg(v + (size_t)y * HWo);
            const float *row = pl + (size_t)y * W;
            for (int x = 0; x < F; ++x) {
                const float hx = __ldg(hz + (size_t)x * HWo);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    if (c0 + cc < C) acc[cc] += __ldg(row + (size_t)cc * HW + x) * vy * hx;   // temp1*temp2*temp3 (:76)
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            if (c0 + cc < C) out[((size_t)b * C + c0 + cc) * HWo + po] = acc[cc];
    }
}

// Tiled forward.  At F = 51 the operator is 3*C*F*F flops per pixel against ~430 bytes: the generic kernel above is
// bound by its loads (one image and one hz load per multiply-add), not by HBM.  Here a CTA owns a 64 x 16 tile of
// output pixels; the (64 + F - 1) x (16 + F - 1) image region of three channels is staged in shared memory once, each
// thread owns a 2 x 2 block of pixels, and the horizontal taps of the four are held in registers 8 at a time: one
// 8-byte shared-memory load feeds eight multiply-adds (two pixels beside, two pixels below -- the lower pair meets
// image row r with its vertical tap r - 1).  The weight v*hz of a term is formed once and shared by the three
// channels; the terms of one pixel are summed chunk by chunk of 8 columns instead of row by row (:65-77).
namespace sct {
constexpr int TWO = 64, THO = 16, NT = 256, XC = 8;
}

template <int XC, bool FULL>
__device__ __forceinline__ void sepconv_chunk(const float *__restrict__ s_img, const float *__restrict__ v,
                                              const float *__restrict__ hz, const size_t (&po)[4], size_t HWo,
                                              int F, int RH, int pitch, int row0, int col0, int xc, int nk, int nc,
                                              float (&acc)[4][3])
{
    float h[4][XC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < XC; ++k) h[q][k] = (FULL || k < nk) ? __ldg(hz + (size_t)(xc + k) * HWo + po[q]) : 0.0f;
    // vertical taps are requested one image row ahead (two CTAs of 8 warps per SM do not hide an L2 round trip)
    float vn[4] = {__ldg(v + po[0]), __ldg(v + po[1]), 0.0f, 0.0f};
    for (int r = 0; r <= F; ++r) {   // image row row0 + r: vertical tap r of the upper pixels, r - 1 of the lower ones
        float vy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) vy[q] = vn[q];
        vn[0] = r + 1 < F ? __ldg(v + (size_t)(r + 1) * HWo + po[0]) : 0.0f;
        vn[1] = r + 1 < F ? __ldg(v + (size_t)(r + 1) * HWo + po[1]) : 0.0f;
        vn[2] = r < F ? __ldg(v + (size_t)r * HWo + po[2]) : 0.0f;
        vn[3] = r < F ? __ldg(v + (size_t)r * HWo + po[3]) : 0.0f;
        float I[3][XC + 2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const float2 *row = reinterpret_cast<const float2 *>(s_img + ((cc < nc ? cc : 0) * RH + row0 + r) * pitch + col0 + xc);
#pragma unroll
            for (int k = 0; k < XC / 2 + 1; ++k) {
                const float2 t = row[k];
                I[cc][2 * k] = t.x; I[cc][2 * k + 1] = t.y;
            }
        }
#pragma unroll
        for (int k = 0; k < XC; ++k) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float wq = vy[q] * h[q][k];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) acc[q][cc] = fmaf(I[cc][k + (q & 1)], wq, acc[q][cc]);
            }
        }
    }
}

__global__ void __launch_bounds__(sct::NT, 2)
sepconv_forward_tiled_kernel(const float *__restrict__ in1, const float *__restrict__ in2,
                             const float *__restrict__ in3, float *__restrict__ out, int C, int H, int W, int F,
                             int pitch)
{
    using namespace sct;
    extern __shared__ __align__(16) float s_img[];   // [3][THO + F - 1][pitch]; columns past the region are zero
    const int Ho = H - F + 1, Wo = W - F + 1, RH = THO + F - 1, RWd = TWO + F - 1;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int w0 = blockIdx.x * TWO, h0 = blockIdx.y * THO, b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    bool valid[4];
    size_t po[4];   // pixel q = 2 * dy + dx; pixels outside the map read pixel 0 of the map and are never stored
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int h_i = h0 + 2 * ty + (q >> 1), w_i = w0 + 2 * lane + (q & 1);
        valid[q] = h_i < Ho && w_i < Wo;
        po[q] = valid[q] ? (size_t)h_i * Wo + w_i : 0;
    }
    const float *v = in2 + (size_t)b * F * HWo, *hz = in3 + (size_t)b * F * HWo;

    for (int c0 = 0; c0 < C; c0 += 3) {
        const int nc = min(3, C - c0);
        __syncthreads();   // the previous pass is done with the region
        for (int row = ty; row < nc * RH; row += NT / 32) {
            const int cc = row / RH, gy = h0 + row - cc * RH;
            const float *src = in1 + ((size_t)b * C + c0 + cc) * HW + (size_t)gy * W;
            for (int rx = lane; rx < pitch; rx += 32)
                s_img[row * pitch + rx] = (rx < RWd && gy < H && w0 + rx < W) ? __ldg(src + w0 + rx) : 0.0f;
        }
        __syncthreads();

        float acc[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 0.0f;
        int xc = 0;
        for (; xc + XC <= F; xc += XC)
            sepconv_chunk<XC, true>(s_img, v, hz, po, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, XC, nc, acc);
        if (F - xc > XC / 2) sepconv_chunk<XC, false>(s_img, v, hz, po, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, F - xc, nc, acc);
        else if (xc < F) sepconv_chunk<XC / 2, false>(s_img, v, hz, po, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, F - xc, nc, acc);   // short tail: half the columns
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                if (cc < nc && valid[q]) out[((size_t)b * C + c0 + cc) * HWo + po[q]] = acc[q][cc];
    }
}

// gradinput2[y] = sum_c g_c sum_x I*hz[x]   (:124-125);  gradinput3[x] = sum_c g_c sum_y I*v[y]   (:126-127)
__global__ void __launch_bounds__(BX *BY)
sepconv_backward_filters_kernel(const float *__restrict__ in1, const float *__restrict__ in2,
                                const float *__restrict__ in3, const float *__restrict__ gout,
                                float *__restrict__ gi2, float *__restrict__ gi3, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= Wo || h_i >= Ho) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo, po = (size_t)h_i * Wo + w_i;
    const float *v = in2 + (size_t)b * F * HWo + po, *hz = in3 + (size_t)b * F * HWo + po;
    const float *g = gout + (size_t)b * C * HWo + po;
    const float *pl = in1 + (size_t)b * C * HW + (size_t)h_i * W + w_i;
    for (int y = 0; y < F; ++y) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gc = __ldg(g + (size_t)c * HWo);
            const float *row = pl + (size_t)c * HW + (size_t)y * W;
            float t = 0.f;
            for (int x = 0; x < F; ++x) t += __ldg(row + x) * __ldg(hz + (size_t)x * HWo);
            s += gc * t;
        }
        gi2[((size_t)b * F + y) * HWo + po] = s;
    }
    for (int x = 0; x < F; ++x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gc = __ldg(g + (size_t)c * HWo);
            const float *col = pl + (size_t)c * HW + x;
            float t = 0.f;
            for (int y = 0; y < F; ++y) t += __ldg(col + (size_t)y * W) * __ldg(v + (size_t)y * HWo);
            s += gc * t;
        }
        gi3[((size_t)b * F + x) * HWo + po] = s;
    }
}

// Tiled filter gradients: the staging and the 2 x 2 pixel blocks of the tiled forward, eight columns at a time.  With
// J[y,x] = sum_c g_c * I_c[h+y, w+x] (three multiply-adds per term), gradinput3[x] = sum_y J * v[y] completes inside a
// column chunk and is stored once; gradinput2[y] = sum_x J * hz[x] is carried from chunk to chunk through the output
// array itself (a thread-private address: plain load / add / store, the old value requested before the row's
// arithmetic).  No atomics, every element written; the reference issues 2*C*F*F atomicAdds per pixel (:124-127).
namespace scb {
constexpr int XC = 8, XH = 4;   // columns per chunk (taps held in registers); columns per shared-memory step
}

template <int XC, bool FULL>
__device__ __forceinline__ void sepconv_filtergrad_chunk(const float *__restrict__ s_img, const float *__restrict__ v,
                                                         const float *__restrict__ hz, float *__restrict__ gi2,
                                                         float *__restrict__ gi3, const size_t (&po)[4],
                                                         const bool (&valid)[4], size_t HWo, int F, int RH, int pitch,
                                                         int row0, int col0, int xc, int nk, const float (&g)[4][3],
                                                         bool first_chunk, bool first_pass)
{
    constexpr int XH = scb::XH;
    float h[4][XC], g3[4][XC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < XC; ++k) {
            h[q][k] = (FULL || k < nk) ? __ldg(hz + (size_t)(xc + k) * HWo + po[q]) : 0.0f;
            g3[q][k] = 0.0f;
        }
    // vertical taps and the carried gradinput2 sums are requested one image row ahead
    float vn[4], on[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool lv = valid[q] && (q >> 1) == 0;
        vn[q] = lv ? __ldg(v + po[q]) : 0.0f;
        on[q] = (lv && !first_chunk) ? gi2[po[q]] : 0.0f;
    }
    for (int r = 0; r <= F; ++r) {   // image row row0 + r: vertical tap r of the upper pixels, r - 1 of the lower ones
        float vy[4], old[4];
        bool live[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int y = r - (q >> 1);
            live[q] = valid[q] && y >= 0 && y < F;
            vy[q] = vn[q];
            old[q] = on[q];
            const bool lvn = valid[q] && y + 1 >= 0 && y + 1 < F;
            const size_t at = (size_t)(lvn ? y + 1 : 0) * HWo + po[q];
            vn[q] = lvn ? __ldg(v + at) : 0.0f;
            on[q] = (lvn && !first_chunk) ? gi2[at] : 0.0f;
        }
        float s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k0 = 0; k0 < XC; k0 += XH) {
            float I[3][XH + 2];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const float2 *row = reinterpret_cast<const float2 *>(s_img + (cc * RH + row0 + r) * pitch + col0 + xc + k0);
#pragma unroll
                for (int k = 0; k < XH / 2 + 1; ++k) {
                    const float2 t = row[k];
                    I[cc][2 * k] = t.x; I[cc][2 * k + 1] = t.y;
                }
            }
#pragma unroll
            for (int k = 0; k < XH; ++k) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = k + (q & 1);
                    const float J = fmaf(g[q][2], I[2][i], fmaf(g[q][1], I[1][i], g[q][0] * I[0][i]));
                    s2[q] = fmaf(J, h[q][k0 + k], s2[q]);
                    g3[q][k0 + k] = fmaf(J, vy[q], g3[q][k0 + k]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (live[q]) gi2[(size_t)(r - (q >> 1)) * HWo + po[q]] = old[q] + s2[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < XC; ++k)
            if (valid[q] && (FULL || k < nk)) {
                float *d = gi3 + (size_t)(xc + k) * HWo + po[q];
                *d = first_pass ? g3[q][k] : *d + g3[q][k];
            }
}

__global__ void __launch_bounds__(sct::NT, 2)
sepconv_backward_filters_tiled_kernel(const float *__restrict__ in1, const float *__restrict__ in2,
                                      const float *__restrict__ in3, const float *__restrict__ gout,
                                      float *__restrict__ gi2, float *__restrict__ gi3, int C, int H, int W, int F,
                                      int pitch)
{
    using namespace sct;
    extern __shared__ __align__(16) float s_img[];   // [3][THO + F - 1][pitch]; zero where there is no image / channel
    const int Ho = H - F + 1, Wo = W - F + 1, RH = THO + F - 1, RWd = TWO + F - 1;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int w0 = blockIdx.x * TWO, h0 = blockIdx.y * THO, b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    bool valid[4];
    size_t po[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int h_i = h0 + 2 * ty + (q >> 1), w_i = w0 + 2 * lane + (q & 1);
        valid[q] = h_i < Ho && w_i < Wo;
        po[q] = valid[q] ? (size_t)h_i * Wo + w_i : 0;
    }
    const float *v = in2 + (size_t)b * F * HWo, *hz = in3 + (size_t)b * F * HWo;
    float *o2 = gi2 + (size_t)b * F * HWo, *o3 = gi3 + (size_t)b * F * HWo;

    for (int c0 = 0; c0 < C; c0 += 3) {
        const int nc = min(3, C - c0);
        __syncthreads();
        for (int row = ty; row < 3 * RH; row += NT / 32) {
            const int cc = row / RH, gy = h0 + row - cc * RH;
            const float *src = in1 + ((size_t)b * C + c0 + (cc < nc ? cc : 0)) * HW + (size_t)gy * W;
            for (int rx = lane; rx < pitch; rx += 32)
                s_img[row * pitch + rx] = (cc < nc && rx < RWd && gy < H && w0 + rx < W) ? __ldg(src + w0 + rx) : 0.0f;
        }
        float g[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                g[q][cc] = (valid[q] && cc < nc) ? __ldg(gout + ((size_t)b * C + c0 + cc) * HWo + po[q]) : 0.0f;
        __syncthreads();

        int xc = 0;
        for (; xc + scb::XC <= F; xc += scb::XC)
            sepconv_filtergrad_chunk<scb::XC, true>(s_img, v, hz, o2, o3, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc,
                                           scb::XC, g, c0 == 0 && xc == 0, c0 == 0);
        if (F - xc > scb::XC / 2)
            sepconv_filtergrad_chunk<scb::XC, false>(s_img, v, hz, o2, o3, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc,
                                            F - xc, g, c0 == 0 && xc == 0, c0 == 0);
        else if (xc < F)
            sepconv_filtergrad_chunk<scb::XC / 2, false>(s_img, v, hz, o2, o3, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc,
                                            F - xc, g, c0 == 0 && xc == 0, c0 == 0);
    }
}

// gradinput1[c,Y,X] = sum over (y,x) with (Y-y, X-x) a valid output pixel of g[c]*v[y]*hz[x] there (:122-123)
__global__ void __launch_bounds__(BX *BY)
sepconv_backward_image_kernel(const float *__restrict__ in2, const float *__restrict__ in3,
                              const float *__restrict__ gout, float *__restrict__ gi1, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int X = blockIdx.x * BX + threadIdx.x, Y = blockIdx.y * BY + threadIdx.y;
    if (X >= W || Y >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    const int y_lo = max(0, Y - (Ho - 1)), y_hi = min(F - 1, Y);
    const int x_lo = max(0, X - (Wo - 1)), x_hi = min(F - 1, X);
    for (int c0 = 0; c0 < C; c0 += 3) {
        float acc[3] = {0.f, 0.f, 0.f};
        for (int y = y_lo; y <= y_hi; ++y) {
            const size_t rowo = (size_t)(Y - y) * Wo;
            for (int x = x_lo; x <= x_hi; ++x) {
                const size_t po = rowo + (X - x);
                const float k = __ldg(in2 + ((size_t)b * F + y) * HWo + po) * __ldg(in3 + ((size_t)b * F + x) * HWo + po);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    if (c0 + cc < C) acc[cc] += __ldg(gout + ((size_t)b * C + c0 + cc) * HWo + po) * k;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            if (c0 + cc < C) gi1[((size_t)b * C + c0 + cc) * HW + (size_t)Y * W + X] = acc[cc];
    }
}

// Tiled image gradient: the transpose of the tiled forward.  A CTA owns a 64 x 16 tile of OUTPUT pixels and sums what
// they send to the (64 + F - 1) x (16 + F - 1) image region in shared memory: per image row r and chunk of 8 columns a
// thread first combines the terms g_c * v * hz of its 2 x 2 pixels in registers (10 columns x 3 channels), then adds
// them to the region with five 8-byte read-modify-writes per channel.  Within a warp instruction the lanes own
// distinct 8-byte slots (a slot is revisited by the neighbouring lane one step later, after a __syncwarp);
// warps own distinct rows as long as they work on the same r, which the barrier per row guarantees.  The region is
// then added to gradinput1 with one atomic per element: ~9 tiles overlap on an image pixel, against the reference's
// F*F atomics per channel and pixel (:122-123).  gradinput1 is zeroed by the launcher.
template <int XC, bool FULL>
__device__ __forceinline__ void sepconv_imagegrad_chunk(float *__restrict__ s_acc, const float *__restrict__ v,
                                                        const float *__restrict__ hz, const size_t (&po)[4],
                                                        const bool (&valid)[4], size_t HWo, int F, int RH, int pitch,
                                                        int row0, int col0, int xc, int nk, const float (&g)[4][3])
{
    float h[4][XC];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < XC; ++k) h[q][k] = (FULL || k < nk) ? __ldg(hz + (size_t)(xc + k) * HWo + po[q]) : 0.0f;
    auto tap = [&](int q, int r) {   // vertical tap of pixel q that meets image row row0 + r (0 when there is none)
        const int y = r - (q >> 1);
        const bool live = valid[q] && y >= 0 && y < F;
        return live ? __ldg(v + (size_t)(live ? y : 0) * HWo + po[q]) : 0.0f;
    };
    float vn[4];   // requested one image row ahead
#pragma unroll
    for (int q = 0; q < 4; ++q) vn[q] = tap(q, 0);
    for (int r = 0; r <= F; ++r) {
        float vy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            vy[q] = vn[q];
            vn[q] = tap(q, r + 1);
        }
        float col[3][XC + 2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int i = 0; i < XC + 2; ++i) col[cc][i] = 0.0f;
#pragma unroll
        for (int k = 0; k < XC; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float wq = vy[q] * h[q][k];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) col[cc][k + (q & 1)] = fmaf(g[q][cc], wq, col[cc][k + (q & 1)]);
            }
        __syncthreads();   // every warp is on row r: the rows row0 + r of different warps are distinct
        float2 *row = reinterpret_cast<float2 *>(s_acc + (row0 + r) * pitch + col0 + xc);
        const int plane = RH * pitch / 2;   // pitch is even
#pragma unroll
        for (int i = 0; i < XC / 2 + 1; ++i) {
            float2 t[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) t[cc] = row[cc * plane + i];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                t[cc].x += col[cc][2 * i]; t[cc].y += col[cc][2 * i + 1];
                row[cc * plane + i] = t[cc];
            }
            __syncwarp();   // slot i of this lane is slot i + 1 of the lane before it: keep the steps in order
        }
    }
}

__global__ void __launch_bounds__(sct::NT, 2)
sepconv_backward_image_tiled_kernel(const float *__restrict__ in2, const float *__restrict__ in3,
                                    const float *__restrict__ gout, float *__restrict__ gi1, int C, int H, int W,
                                    int F, int pitch)
{
    using namespace sct;
    extern __shared__ __align__(16) float s_acc[];   // [3][THO + F - 1][pitch]
    const int Ho = H - F + 1, Wo = W - F + 1, RH = THO + F - 1, RWd = TWO + F - 1;
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int w0 = blockIdx.x * TWO, h0 = blockIdx.y * THO, b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    bool valid[4];
    size_t po[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int h_i = h0 + 2 * ty + (q >> 1), w_i = w0 + 2 * lane + (q & 1);
        valid[q] = h_i < Ho && w_i < Wo;
        po[q] = valid[q] ? (size_t)h_i * Wo + w_i : 0;
    }
    const float *v = in2 + (size_t)b * F * HWo, *hz = in3 + (size_t)b * F * HWo;

    for (int c0 = 0; c0 < C; c0 += 3) {
        const int nc = min(3, C - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * RH * pitch; i += NT) s_acc[i] = 0.0f;
        float g[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                g[q][cc] = (valid[q] && cc < nc) ? __ldg(gout + ((size_t)b * C + c0 + cc) * HWo + po[q]) : 0.0f;
        __syncthreads();

        int xc = 0;
        for (; xc + XC <= F; xc += XC)
            sepconv_imagegrad_chunk<XC, true>(s_acc, v, hz, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, XC, g);
        if (F - xc > XC / 2)
            sepconv_imagegrad_chunk<XC, false>(s_acc, v, hz, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, F - xc, g);
        else if (xc < F)
            sepconv_imagegrad_chunk<XC / 2, false>(s_acc, v, hz, po, valid, HWo, F, RH, pitch, 2 * ty, 2 * lane, xc, F - xc, g);
        __syncthreads();
        for (int row = ty; row < nc * RH; row += NT / 32) {
            const int cc = row / RH, gy = h0 + row - cc * RH;
            if (gy >= H) continue;
            float *dst = gi1 + ((size_t)b * C + c0 + cc) * HW + (size_t)gy * W + w0;
            for (int rx = lane; rx < RWd && w0 + rx < W; rx += 32) atomicAdd(dst + rx, s_acc[row * pitch + rx]);
        }
    }
}

// SeparableConvFlow: flow = centroid of the weights - (F-1)/2, sentinel -2000 when |sum| == 0 (:56-89)
__global__ void __launch_bounds__(256)
sepconvflow_forward_kernel(const float *__restrict__ in2, const float *__restrict__ in3, float *__restrict__ flow,
                           size_t HWo, size_t total, int F)
{
    // one thread per (pixel, flow channel); the grid covers the map exactly (no grid-stride tail)
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = blockIdx.y;   // channel 0 <- input3 (x), channel 1 <- input2 (y)
    const size_t b = idx / HWo, po = idx - b * HWo;
    const float *src = (ch == 0 ? in3 : in2) + b * F * HWo + po;
    float num = 0.f, den = 0.f;
    int k = 0;
    for (; k + 8 <= F; k += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = ld_stream(src + (size_t)(k + j) * HWo);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            num += (float)(k + j) * t[j];   // :61
            den += t[j];                    // :62
        }
    }
    for (; k < F; ++k) {
        const float t = ld_stream(src + (size_t)k * HWo);
        num += (float)k * t;
        den += t;
    }
    // flow_y / sum_weights - ((float)(filter_size)-1.0)/2.0 is evaluated in double (:66)
    const float val = (float)((double)(num / den) - ((double)(float)F - 1.0) / 2.0);
    st_stream(flow + (b * 2 + ch) * HWo + po, fabsf(den) > 0.0f ? val : -2000.0f);
}

// gi[k] = g * (k / S - num / S^2) where |S| > 0, else 0 (:131-169)
__global__ void __launch_bounds__(256)
sepconvflow_backward_kernel(const float *__restrict__ in2, const float *__restrict__ in3,
                            const float *__restrict__ gflow, float *__restrict__ gi2, float *__restrict__ gi3,
                            size_t HWo, size_t total, int F)
{
    // one thread per (pixel, flow channel); the grid covers the map exactly
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ch = blockIdx.y;
    const size_t b = idx / HWo, po = idx - b * HWo;
    const float *src = (ch == 0 ? in3 : in2) + b * F * HWo + po;
    float *dst = (ch == 0 ? gi3 : gi2) + b * F * HWo + po;
    const float g = ld_stream(gflow + (b * 2 + ch) * HWo + po);
    float num = 0.f, den = 0.f;
    int k = 0;
    for (; k + 8 <= F; k += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = ld_stream(src + (size_t)(k + j) * HWo);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            num += (float)(k + j) * t[j];
            den += t[j];
        }
    }
    for (; k < F; ++k) {
        const float t = ld_stream(src + (size_t)k * HWo);
        num += (float)k * t;
        den += t;
    }
    if (fabsf(den) > 0.0f) {
        const float offset = num / (den * den);   // :138
        for (int k2 = 0; k2 < F; ++k2) st_stream(dst + (size_t)k2 * HWo, g * ((float)k2 / den - offset));   // :140-143
    } else {
        for (int k2 = 0; k2 < F; ++k2) st_stream(dst + (size_t)k2 * HWo, 0.0f);
    }
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_separableconv_forward(const float *input1, const float *input2, const float *input3,
                                            float *output, int B, int C, int H, int W, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || F <= 0 || H - F + 1 <= 0 || W - F + 1 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !input3 || !output) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    {   // tiled kernel whenever its image region fits in shared memory (F <= ~110)
        const int pitch = (sct::TWO + F - 1 + sct::XC + 1) & ~1;
        const size_t smem = (size_t)3 * (sct::THO + F - 1) * pitch * sizeof(float);
        if (smem <= 112 * 1024 && ceil_div(H - F + 1, sct::THO) <= 65535u &&
            cudaFuncSetAttribute(sepconv_forward_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
            dim3 grid(ceil_div(W - F + 1, sct::TWO), ceil_div(H - F + 1, sct::THO), B);
            sepconv_forward_tiled_kernel<<<grid, sct::NT, smem, (cudaStream_t)stream>>>(input1, input2, input3, output, C, H, W, F, pitch);
            note_launch();
            return check_launch("separableconv forward");
        }
        (void)cudaGetLastError();
    }
    dim3 block(BX, BY), grid(ceil_div(W - F + 1, BX), ceil_div(H - F + 1, BY), B);
    sepconv_forward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(input1, input2, input3, output, C, H, W, F);
    note_launch();
    return check_launch("separableconv forward");
}

VFIDKR_API int vfidkr_separableconv_backward(const float *input1, const float *input2, const float *input3,
                                             const float *gradoutput, float *gradinput1, float *gradinput2,
                                             float *gradinput3, int B, int C, int H, int W, int F,
                                             vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || F <= 0 || H - F + 1 <= 0 || W - F + 1 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !input3 || !gradoutput || !gradinput1 || !gradinput2 || !gradinput3) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(BX, BY), grid_o(ceil_div(W - F + 1, BX), ceil_div(H - F + 1, BY), B), grid_i(ceil_div(W, BX), ceil_div(H, BY), B);
    const int pitch = (sct::TWO + F - 1 + sct::XC + 1) & ~1;
    const size_t smem = (size_t)3 * (sct::THO + F - 1) * pitch * sizeof(float);
    if (smem <= 112 * 1024 && ceil_div(H - F + 1, sct::THO) <= 65535u &&
        cudaFuncSetAttribute(sepconv_backward_filters_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        dim3 grid_t(ceil_div(W - F + 1, sct::TWO), ceil_div(H - F + 1, sct::THO), B);
        sepconv_backward_filters_tiled_kernel<<<grid_t, sct::NT, smem, s>>>(input1, input2, input3, gradoutput, gradinput2, gradinput3, C, H, W, F, pitch);
    } else {
        (void)cudaGetLastError();
        sepconv_backward_filters_kernel<<<grid_o, block, 0, s>>>(input1, input2, input3, gradoutput, gradinput2, gradinput3, C, H, W, F);
    }
    if (smem <= 112 * 1024 && ceil_div(H - F + 1, sct::THO) <= 65535u &&
        cudaFuncSetAttribute(sepconv_backward_image_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        dim3 grid_t(ceil_div(W - F + 1, sct::TWO), ceil_div(H - F + 1, sct::THO), B);
        cudaMemsetAsync(gradinput1, 0, (size_t)B * C * H * W * sizeof(float), s);
        sepconv_backward_image_tiled_kernel<<<grid_t, sct::NT, smem, s>>>(input2, input3, gradoutput, gradinput1, C, H, W, F, pitch);
    } else {
        (void)cudaGetLastError();
        sepconv_backward_image_kernel<<<grid_i, block, 0, s>>>(input2, input3, gradoutput, gradinput1, C, H, W, F);
    }
    note_launch(2);
    return check_launch("separableconv backward");
}

VFIDKR_API int vfidkr_separableconvflow_forward(const float *input2, const float *input3, float *flow_output,
                                                int B, int Ho, int Wo, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || Ho <= 0 || Wo <= 0 || F <= 0 || !input2 || !input3 || !flow_output) return VFIDKR_ERR_ARG;
    const size_t HWo = (size_t)Ho * Wo, total = (size_t)B * HWo;
    if ((total + 255) / 256 > 0x7fffffffull) return VFIDKR_ERR_ARG;
    sepconvflow_forward_kernel<<<dim3((unsigned)((total + 255) / 256), 2), 256, 0, (cudaStream_t)stream>>>(input2, input3, flow_output, HWo, total, F);
    note_launch();
    return check_launch("separableconvflow forward");
}

VFIDKR_API int vfidkr_separableconvflow_backward(const float *input2, const float *input3,
                                                 const float *gradflow_output, float *gradinput2, float *gradinput3,
                                                 int B, int Ho, int Wo, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || Ho <= 0 || Wo <= 0 || F <= 0) return VFIDKR_ERR_ARG;
    if (!input2 || !input3 || !gradflow_output || !gradinput2 || !gradinput3) return VFIDKR_ERR_ARG;
    const size_t HWo = (size_t)Ho * Wo, total = (size_t)B * HWo;
    if ((total + 255) / 256 > 0x7fffffffull) return VFIDKR_ERR_ARG;
    sepconvflow_backward_kernel<<<dim3((unsigned)((total + 255) / 256), 2), 256, 0, (cudaStream_t)stream>>>(input2, input3, gradflow_output, gradinput2, gradinput3, HWo, total, F);
    note_launch();
    return check_launch("separableconvflow backward");
}
