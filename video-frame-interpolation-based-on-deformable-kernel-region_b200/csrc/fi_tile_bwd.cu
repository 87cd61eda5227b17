// fi_tile_bwd.cu -- FilterInterpolation backward (all four families, F = 4) with warp-private shared-memory
// accumulation of the image gradient, sm_100a.
//
// What is computed follows my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu
//   :2827-3125 "_ori"   :430-1215 4-input DKR   :1500-1935 "_deforconv"   :2195-2567 "_nofilterwithdeforconv".
// The reference (and the per-pixel kernel in filterinterpolation.cu) sends 16*C scalar REDs per pixel to the image
// gradient; that kernel is latency-bound behind the L2's atomic unit (ncu: nothing above 32 % of peak, 15 warps per
// issue stalled on the scoreboard).  Shared-memory float atomics are no way out -- sm_100a has no native shared
// fp32 add, atomicAdd compiles to an ATOMS.CAST.SPIN loop that retries heavily when eight warps share a tile
// (measured: 4100 instructions per warp, slower than the REDs).  So the accumulation is made contention-free by
// construction instead:
//   * a WARP owns a 32 x 4 pixel tile and a private region of shared memory (768 cells per channel) that holds the
//     bounding box of its pixels' 4 x 4 windows; it walks its four pixel rows one after the other;
//   * within one row all lanes execute the same tap at the same time, so two lanes collide iff they have the same
//     window origin (L, T): one __match_any per pixel row ranks such duplicates and the read-modify-write of a tap is
//     issued once per rank (one pass when the origins are distinct -- the common case), plain LDS / FADD / STS,
//     __syncwarp between taps because neighbouring lanes' windows overlap;
//   * pixels whose window is clamped at the image border (clamping merges taps of different lanes) and tiles whose
//     box does not fit the region go straight to global REDs -- correctness never depends on the flow;
//   * at the end the warp flushes its box with coalesced REDs, skipping cells that received nothing: about 0.4
//     warp-level REDs per pixel instead of 48;
//   * flow / filter / offset gradients are thread-private: register accumulation, plain streaming stores.
// No block-level synchronisation exists; the CTA only packs eight warp tiles.  gradinput1 must be zero on entry
// (the launcher clears it on the stream).
#include <climits>

#include "common.cuh"
#include "fi_common.cuh"

namespace vfidkr {
namespace {

namespace tb {
constexpr int TW = 32, ROWS = 4, WARPS = 8, NT = TW * WARPS;
constexpr int CELLS = 768;   // cells per channel of one warp's private gradient region
}  // namespace tb

template <int V, int CCH>
__global__ void __launch_bounds__(tb::NT, 3)
fi_backward_warptile_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                            const float *__restrict__ in4, const float *__restrict__ gout, float *__restrict__ gi1,
                            float *__restrict__ gi2, float *__restrict__ gi3, float *__restrict__ gi4,
                            int C, int H, int W)
{
    using namespace tb;
    constexpr int F = 4, T2 = 16;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x, warp = threadIdx.y;
    float *G = smem + warp * (CCH * CELLS);   // [CCH][bh][pitch], this warp only

    const int w_i = blockIdx.x * TW + lane;
    const int hy0 = (blockIdx.y * WARPS + warp) * ROWS;
    if (hy0 >= H) return;   // whole warp outside (no block-level barrier anywhere below)
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;

    // ---- flow of the warp's 4 rows; bounding box of the windows that will use the shared region ----
    float fxr[ROWS], fyr[ROWS];
    int bx0 = INT_MAX, by0 = INT_MAX, bx1 = INT_MIN, by1 = INT_MIN;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int h_i = hy0 + r;
        fxr[r] = 0.0f; fyr[r] = 0.0f;
        if (w_i < W && h_i < H) {
            const size_t pix = (size_t)h_i * W + w_i;
            fxr[r] = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
            fyr[r] = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int h_i = hy0 + r;
        const FiPix p = fi_pixel(w_i, h_i, fxr[r], fyr[r], W, H, F);
        if (w_i < W && h_i < H && p.in_range && p.L >= 0 && p.T >= 0 && p.L + 3 < W && p.T + 3 < H) {
            bx0 = min(bx0, p.L); bx1 = max(bx1, p.L + 3);
            by0 = min(by0, p.T); by1 = max(by1, p.T + 3);
        }
    }
    bx0 = __reduce_min_sync(FULL, bx0); by0 = __reduce_min_sync(FULL, by0);
    bx1 = __reduce_max_sync(FULL, bx1); by1 = __reduce_max_sync(FULL, by1);
    const bool any = bx1 >= bx0;
    const int bw = any ? bx1 - bx0 + 1 : 0, bh = any ? by1 - by0 + 1 : 0;
    // pitch 64 keeps a warp whose lanes fall on two window rows bank-conflict free; otherwise as tight as the box
    const int pitch = (bw <= 64 && bh * 64 <= CELLS) ? 64 : ((bw + 7) & ~7);
    const bool fits = any && pitch * bh <= CELLS;   // warp-uniform

    for (int c0 = 0; c0 < C; c0 += CCH) {
        const float *img = in1 + ((size_t)b * C + c0) * HW;
        float *gimg = gi1 + ((size_t)b * C + c0) * HW;
        const int nc = min(CCH, C - c0);
        const bool first = (c0 == 0);
        if (fits) {
            for (int idx = lane * 4; idx < CCH * CELLS; idx += 128) *reinterpret_cast<float4 *>(G + idx) = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
        }
#pragma unroll 1
        for (int r = 0; r < ROWS; ++r) {
            const int h_i = hy0 + r;
            const bool inside = w_i < W && h_i < H;
            const size_t pix = inside ? (size_t)h_i * W + w_i : 0;
            const FiPix p = fi_pixel(w_i, h_i, fxr[r], fyr[r], W, H, F);
            const bool active = inside && p.in_range;
            const bool border = !(p.L >= 0 && p.T >= 0 && p.L + 3 < W && p.T + 3 < H);
            const bool use_s = fits && active && !border;
            // lanes with the same window origin collide on every tap: rank them (warp-uniform control flow here)
            const unsigned peers = __match_any_sync(FULL, use_s ? p.T * W + p.L : (int)(0x80000000u | (unsigned)lane));
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int maxrank = __reduce_max_sync(FULL, rank);

            float *g2 = gi2 + (size_t)b * 2 * HW + pix;
            float *g3 = (V == V_NOFILT) ? nullptr : gi3 + (size_t)b * T2 * HW + pix;
            float *go = (V == V_ORI) ? nullptr : (V == V_NOFILT ? gi3 : gi4) + (size_t)b * 2 * T2 * HW + pix;
            const float *wp = (V == V_NOFILT) ? nullptr : in3 + (size_t)b * T2 * HW + pix;
            const float *op = (V == V_ORI) ? nullptr : (V == V_NOFILT ? in3 : in4) + (size_t)b * 2 * T2 * HW + pix;

            if (inside && !p.in_range && first) {   // contributes nothing; the reference leaves the caller's zeros (:2863)
                st_stream(g2, 0.0f);
                st_stream(g2 + HW, 0.0f);
#pragma unroll
                for (int k = 0; k < T2; ++k) {
                    if (V != V_NOFILT) st_stream(g3 + (size_t)k * HW, 0.0f);
                    if (V != V_ORI) { st_stream(go + (size_t)k * HW, 0.0f); st_stream(go + (size_t)(T2 + k) * HW, 0.0f); }
                }
            }
            float g[CCH];
#pragma unroll
            for (int cc = 0; cc < CCH; ++cc)
                g[cc] = (active && cc < nc) ? ld_stream(gout + ((size_t)b * C + c0 + cc) * HW + pix) : 0.0f;
            float gx = 0.0f, gy = 0.0f;
            const int sbase = (p.T - by0) * pitch + (p.L - bx0);   // region cell of tap (0, 0); meaningful when use_s

#pragma unroll 1
            for (int j = 0; j < F; ++j) {
                const int cy = clampi(p.T + j, 0, H - 1);
#pragma unroll
                for (int i = 0; i < F; ++i) {
                    const int k = j * F + i;
                    float add[CCH];
#pragma unroll
                    for (int cc = 0; cc < CCH; ++cc) add[cc] = 0.0f;
                    if (active) {
                        const int cx = clampi(p.L + i, 0, W - 1);
                        const int a = cy * W + cx;   // plane offset of the undeformed tap
                        const float wgt = (V == V_NOFILT) ? 1.0f : ld_stream(wp + (size_t)k * HW);
                        float s3 = 0.0f, soy = 0.0f, sox = 0.0f;
                        if (V == V_ORI) {
                            const QuadCoef qc = quad_coef(j < F / 2, i < F / 2, p.alpha, p.beta);
#pragma unroll
                            for (int cc = 0; cc < CCH; ++cc)
                                if (cc < nc) {
                                    const float gq = g[cc] * qc.q;                  // :2885
                                    const float v = __ldg(img + (size_t)cc * HW + a);
                                    add[cc] = gq * wgt;                             // :2890-2892
                                    s3 += gq * v;                                   // :2893-2895
                                    const float t = g[cc] * (v * wgt);
                                    gx += qc.cx * t;
                                    gy += qc.cy * t;
                                }
                        } else {
                            const float oy = ld_stream(op + (size_t)k * HW), ox = ld_stream(op + (size_t)(T2 + k) * HW);
                            const Deform d = fi_deform(cy, cx, oy, ox, p, H, W);
                            const bool top = (V == V_DKR) ? (j < F / 2) : d.top;
                            const bool left = (V == V_DKR) ? (i < F / 2) : d.left;
                            const QuadCoef qc = quad_coef(top, left, p.alpha, p.beta);
                            const float PTL = (1 - d.phiX) * (1 - d.phiY), PTR = d.phiX * (1 - d.phiY);
                            const float PBL = (1 - d.phiX) * d.phiY, PBR = d.phiY * d.phiX;
#pragma unroll
                            for (int cc = 0; cc < CCH; ++cc)
                                if (cc < nc) {
                                    const float *q = img + (size_t)cc * HW;
                                    const float vTL = __ldg(q + d.aTL), vTR = __ldg(q + d.aTR);
                                    const float vBL = __ldg(q + d.aBL), vBR = __ldg(q + d.aBR);
                                    const float S = PTL * vTL + PTR * vTR + PBL * vBL + PBR * vBR;
                                    const float dSy = -(1 - d.phiX) * vTL + (1 - d.phiX) * vBL - d.phiX * vTR + d.phiX * vBR;  // :986-989
                                    const float dSx = -(1 - d.phiY) * vTL + (1 - d.phiY) * vTR - d.phiY * vBL + d.phiY * vBR;  // :1104-1107
                                    const float gq = g[cc] * qc.q;
                                    add[cc] = gq * wgt;                             // undeformed tap (:497-499, :2258)
                                    s3 += gq * S;                                   // :520-522
                                    soy += gq * dSy * wgt;                          // :990-993
                                    sox += gq * dSx * wgt;                          // :1108-1111
                                    const float t = g[cc] * (S * wgt);
                                    gx += qc.cx * t;
                                    gy += qc.cy * t;
                                }
                        }
                        if (!use_s) {
#pragma unroll
                            for (int cc = 0; cc < CCH; ++cc)
                                if (cc < nc) red_add(gimg + (size_t)cc * HW + a, add[cc]);
                        }
                        // thread-private gradients: plain stores, accumulated across channel chunks in place
                        if (V != V_NOFILT) {
                            float *q3 = g3 + (size_t)k * HW;
                            if (first) st_stream(q3, s3); else *q3 += s3;
                        }
                        if (V != V_ORI) {
                            float *qy = go + (size_t)k * HW, *qx = go + (size_t)(T2 + k) * HW;
                            if (first) { st_stream(qy, soy); st_stream(qx, sox); } else { *qy += soy; *qx += sox; }
                        }
                    }
                    if (fits) {   // warp-uniform: every lane takes part in the hand-over between taps
                        const int sa = sbase + j * pitch + i;
                        for (int round = 0; round <= maxrank; ++round) {
                            if (use_s && rank == round) {
#pragma unroll
                                for (int cc = 0; cc < CCH; ++cc)
                                    if (cc < nc) G[cc * CELLS + sa] += add[cc];
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            if (active) {
                if (first) { st_stream(g2, gx); st_stream(g2 + HW, gy); }
                else { *g2 += gx; *(g2 + HW) += gy; }
            }
        }
        if (fits) {
            // flush the box: one coalesced RED per cell that received something
            for (int cc = 0; cc < nc; ++cc)
                for (int rr = 0; rr < bh; ++rr) {
                    const float *sg = G + cc * CELLS + rr * pitch;
                    float *dst = gimg + (size_t)cc * HW + (size_t)(by0 + rr) * W + bx0;
                    for (int col = lane; col < bw; col += 32) {
                        const float v = sg[col];
                        if (v != 0.0f) red_add(dst + col, v);
                    }
                }
            __syncwarp();
        }
    }
}

template <int V, int CCH>
int launch_tile(const float *in1, const float *in2, const float *in3, const float *in4, const float *gout,
                float *gi1, float *gi2, float *gi3, float *gi4, int B, int C, int H, int W, cudaStream_t s)
{
    using namespace tb;
    constexpr size_t smem = sizeof(float) * WARPS * CCH * CELLS;
    auto kern = fi_backward_warptile_kernel<V, CCH>;
    if (set_error(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tile backward smem"))
        return VFIDKR_ERR_CUDA;
    dim3 block(TW, WARPS), grid(ceil_div(W, TW), ceil_div(H, ROWS * WARPS), B);
    kern<<<grid, block, smem, s>>>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, C, H, W);
    note_launch();
    return check_launch("filterinterpolation backward (warp tiles)");
}

}  // namespace

// F == 4 backward of family `variant` with warp-private shared-memory accumulation; gi1 already cleared by the caller.
int fi_tile_backward(int variant, const float *in1, const float *in2, const float *in3, const float *in4,
                     const float *gout, float *gi1, float *gi2, float *gi3, float *gi4,
                     int B, int C, int H, int W, cudaStream_t s)
{
    if (ceil_div(H, tb::ROWS * tb::WARPS) > 65535u) return -1;
#define VFIDKR_TILE_CASE(VV)                                                                                      \
    case VV:                                                                                                      \
        return C == 3 ? launch_tile<VV, 3>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, B, C, H, W, s)           \
                      : launch_tile<VV, 4>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, B, C, H, W, s);
    switch (variant) {
        VFIDKR_TILE_CASE(V_ORI)
        VFIDKR_TILE_CASE(V_DKR)
        VFIDKR_TILE_CASE(V_DEFOR)
        VFIDKR_TILE_CASE(V_NOFILT)
    }
#undef VFIDKR_TILE_CASE
    return -1;
}

}  // namespace vfidkr
