// filterinterpolation.cu -- adaptive warping with per-pixel F x F filters and its three
// deformable-kernel-region (DKR) variants, forward and backward, for sm_100a.
//
// Behavioural contract (what is computed) follows the reference kernels
//   my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu
//     :2692-3125 "_ori"   :29-1215 4-input DKR   :1353-1935 "_deforconv"   :2070-2567 "_nofilterwithdeforconv"
// The implementation is new:
//   * one thread per output pixel keeps the F*F filter taps (and 2*F*F offsets) in registers
//     and loops channels inside, instead of re-reading them from HBM for every channel;
//   * the backward fuses the reference's 3 (ori) / 5 (DKR) passes over the taps into one and
//     accumulates the thread-private gradients (flow, filter, offsets) in registers with plain
//     stores -- only the image gradient, which really scatters, uses RED atomics.  Two shared-memory
//     accumulation schemes for that scatter were built and measured slower than the REDs (CTA tile
//     with shared atomics: fp32 atomicAdd on shared memory is an ATOMS.CAST.SPIN retry loop, 4100
//     instructions per warp; warp-private tile with ranked duplicates: twice the instructions and
//     the shared-memory carve-out leaves ~30 KB of L1 for the image gathers) -- see DESIGN.md 4.2;
//   * out-of-range pixels write their zeros/copies themselves, so no buffer but gradinput1
//     needs clearing;
//   * 64-bit plane offsets (B*C*H*W exceeds 2^31 for the 196-channel context tensors at 1080p).
#include "common.cuh"
#include "fi_common.cuh"
#include "tma.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>

namespace vfidkr {

// fi_strip.cu / fi_strip_w128.cu: the same kernel with 144- and 128-column tiles; -1 = not applicable
int fi_strip_forward_ori_w144(const float *in1, const float *in2, const float *in3, float *out,
                              int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s);
int fi_strip_forward_ori_w128(const float *in1, const float *in2, const float *in3, float *out,
                              int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s);
// Tile width per launch.  Measured on B200 (profiles/r02/strip_width_ab_v1.log, 128 | 144 columns): 8 x 3 x 1152 x 1984
// 399 | 369 us, blend pair 813 | 790, 8 x 4 x 1152 x 1984 433 | 422, 8 x 3 x 736 x 1280 199 | 185, 16 x 3 x 256 x 448
// 101 | 73, 4K x 2 368 | 356, 4K x 1 201 | 208 -- the 144-column kernel (19 instead of 17 warps per SM at the same
// register budget, never more strips than the 128-column one) wins everywhere except a single 4K frame (3 %), so it is
// taken whenever it applies (W >= 192); the 128-column kernel serves 160 <= W < 192.
static int fi_strip_forward_ori(const float *in1, const float *in2, const float *in3, float *out,
                                int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s)
{
    static const int forced_tw = [] { const char *e = getenv("VFIDKR_FI_STRIP_TW"); return e ? atoi(e) : 0; }();   // 128 / 144: experiments
    if (forced_tw != 128) {
        const int e = fi_strip_forward_ori_w144(in1, in2, in3, out, B, C, H, W, scale, accumulate, out_bs, s);
        if (e >= 0) return e;
    }
    return fi_strip_forward_ori_w128(in1, in2, in3, out, B, C, H, W, scale, accumulate, out_bs, s);
}
int fi_bigc_forward_ori(const float *in1, const float *in2, const float *in3, float *out,
                        int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s);    // fi_bigc.cu (C > 4); -1 = not applicable
int fi_strip_forward_dkr(int variant, const float *in1, const float *in2, const float *filt, const float *offs, float *out,
                         int B, int C, int H, int W, cudaStream_t s);   // fi_strip_dkr.cu; -1 = not applicable

namespace {

constexpr int BX = 32, BY = 8;  // thread block = 32 x 8 output pixels

// The "_ori" forward has three implementations: strip (rolling shared-memory window, fi_strip.cu), tile
// (TMA-streamed taps, gathers through L1) and direct.  The launcher picks the first that applies.  The tests
// check that all of them agree by forcing one through vfidkr_debug_force_forward_path() (a process-wide test
// hook: one relaxed atomic load per launch).  The environment variable VFIDKR_FI_FWD_PATH=strip|tile|direct
// sets its initial value and is read ONCE, at the first launch; never needed in production.
enum { PATH_AUTO = 0, PATH_STRIP, PATH_TILE, PATH_DIRECT };
std::atomic<int> g_forced_path{-1};   // -1: environment not consulted yet
inline int forced_forward_path()
{
    int v = g_forced_path.load(std::memory_order_relaxed);
    if (v >= 0) return v;
    v = PATH_AUTO;
    if (const char *e = env_once("VFIDKR_FI_FWD_PATH"))
        v = e[0] == 's' ? PATH_STRIP : e[0] == 't' ? PATH_TILE : e[0] == 'd' ? PATH_DIRECT : PATH_AUTO;
    int expected = -1;
    g_forced_path.compare_exchange_strong(expected, v, std::memory_order_relaxed);
    return g_forced_path.load(std::memory_order_relaxed);
}

// resident blocks per SM the register allocator must leave room for
constexpr int MINB_FWD_ORI = 4, MINB_FWD_DKR = 3;
__host__ __device__ constexpr int minb_bwd(int V) { return V == V_ORI ? 3 : 2; }

// ------------------------------------------------------------------------------------------
// "_ori" forward.  FT > 0: compile-time filter size, the F*F taps live in registers and the
// channel loop is outermost (the reference re-reads the taps for every channel, :2755).
// FT == 0: run-time filter size, taps re-read through L1.
// ------------------------------------------------------------------------------------------
template <int FT>
__global__ void __launch_bounds__(BX *BY, MINB_FWD_ORI)
fi_forward_ori_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                      float *__restrict__ out, int C, int H, int W, int Frt, float scale, int accumulate, size_t out_bs)
{
    // epilogue: output = scale * result (+ what output held) -- see vfidkr_filterinterpolation_forward_ori_blend
    auto put = [&](float *dst, float v) {
        v *= scale;
        if (accumulate) v += __ldcs(dst);
        st_stream(dst, v);
    };
    const int F = FT > 0 ? FT : Frt;
    const int w_i = blockIdx.x * BX + threadIdx.x;
    const int h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)h_i * W + w_i;

    const float fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
    const FiPix p = fi_pixel(w_i, h_i, fx, fy, W, H, F);

    const float *img = in1 + (size_t)b * C * HW;
    float *o = out + (size_t)b * out_bs + pix;   // out_bs: elements between batch items of the output
    if (!p.in_range) {  // :2814-2819 copies input1
        for (int c = 0; c < C; ++c) put(o + (size_t)c * HW, __ldg(img + (size_t)c * HW + pix));
        return;
    }
    const float *wp = in3 + (size_t)b * F * F * HW + pix;
    const float qTL = (1 - p.alpha) * (1 - p.beta), qTR = p.alpha * (1 - p.beta);
    const float qBL = (1 - p.alpha) * p.beta, qBR = p.alpha * p.beta;

    if (FT > 0) {
        constexpr int FN = FT > 0 ? FT : 1;
        float w[FN * FN];
#pragma unroll
        for (int k = 0; k < FT * FT; ++k) w[k] = ld_stream(wp + (size_t)k * HW);
        int ro[FN], co[FN];
#pragma unroll
        for (int j = 0; j < FT; ++j) {
            ro[j] = clampi(p.T + j, 0, H - 1) * W;   // :2751
            co[j] = clampi(p.L + j, 0, W - 1);       // :2753
        }
        for (int c = 0; c < C; ++c) {
            const float *pl = img + (size_t)c * HW;
            float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < FT; ++j)
#pragma unroll
                for (int i = 0; i < FT; ++i)
                    Q[(j < FT / 2 ? 0 : 2) + (i < FT / 2 ? 0 : 1)] += __ldg(pl + ro[j] + co[i]) * w[j * FT + i];
            put(o + (size_t)c * HW, qTL * Q[0] + qTR * Q[1] + qBL * Q[2] + qBR * Q[3]);
        }
    } else {
        for (int c = 0; c < C; ++c) {
            const float *pl = img + (size_t)c * HW;
            float TL = 0.f, TR = 0.f, BL = 0.f, BR = 0.f;
            for (int j = 0; j < F; ++j) {
                const int r = clampi(p.T + j, 0, H - 1) * W;
                for (int i = 0; i < F; ++i) {
                    const float t = __ldg(pl + r + clampi(p.L + i, 0, W - 1)) * __ldg(wp + (size_t)(j * F + i) * HW);
                    const bool top = j < F / 2, left = i < F / 2;
                    if (top) { if (left) TL += t; else TR += t; } else { if (left) BL += t; else BR += t; }
                }
            }
            put(o + (size_t)c * HW, qTL * TL + qTR * TR + qBL * BL + qBR * BR);
        }
    }
}

// ------------------------------------------------------------------------------------------
// "_ori" forward, TMA-streamed (the production path for F == 4 when W % 4 == 0 and the bases are
// 16-byte aligned).  84 % of the algorithmic bytes of this op are the flow + filter planes, read
// exactly once.  A persistent CTA walks tiles of TW x TH pixels; one elected thread streams the 18
// planes of the NEXT tile into shared memory with two cp.async.bulk.tensor (TMA) box loads while the
// 256 threads work on the current tile, so HBM latency is decoupled from the gather/FMA phase and no
// registers are spent on loads in flight.  Per pixel the 16 taps and the flow are then read from
// shared memory (conflict-free, stride-1) and the image window is gathered through L1/L2.
// ------------------------------------------------------------------------------------------
namespace tmafwd {
constexpr int TW = 64, TH = 4, NTHREADS = TW * TH, STAGES = 2;
constexpr int FILT_FLOATS = 16 * TH * TW, FLOW_FLOATS = 2 * TH * TW;
constexpr uint32_t STAGE_BYTES = (FILT_FLOATS + FLOW_FLOATS) * sizeof(float);
}  // namespace tmafwd

// one channel of one pixel: 4x4 window starting at `win` (row stride W), all taps inside the plane
__device__ __forceinline__ float fi_window_interior(const float *__restrict__ win, int W, const float (&w)[16],
                                                    float qTL, float qTR, float qBL, float qBR)
{
    const float *r0 = win, *r1 = win + W, *r2 = r1 + W, *r3 = r2 + W;
    // the 16 loads carry immediate offsets off four row pointers; quadrant sums in the reference's tap order
    float TL = __ldg(r0) * w[0];
    TL += __ldg(r0 + 1) * w[1];
    float TR = __ldg(r0 + 2) * w[2];
    TR += __ldg(r0 + 3) * w[3];
    TL += __ldg(r1) * w[4];
    TL += __ldg(r1 + 1) * w[5];
    TR += __ldg(r1 + 2) * w[6];
    TR += __ldg(r1 + 3) * w[7];
    float BL = __ldg(r2) * w[8];
    BL += __ldg(r2 + 1) * w[9];
    float BR = __ldg(r2 + 2) * w[10];
    BR += __ldg(r2 + 3) * w[11];
    BL += __ldg(r3) * w[12];
    BL += __ldg(r3 + 1) * w[13];
    BR += __ldg(r3 + 2) * w[14];
    BR += __ldg(r3 + 3) * w[15];
    return qTL * TL + qTR * TR + qBL * BL + qBR * BR;
}

// CT > 0: compile-time channel count (fully unrolled, all gathers independent); CT == 0: run-time C
template <int CT>
__global__ void __launch_bounds__(tmafwd::NTHREADS, 4)
fi_forward_ori_tma_kernel(const __grid_constant__ CUtensorMap map_flow, const __grid_constant__ CUtensorMap map_filt,
                          const float *__restrict__ in1, float *__restrict__ out,
                          int Crt, int H, int W, int tiles_x, int tiles_y, int num_tiles,
                          const FastDiv div_tiles_x, const FastDiv div_tiles_image)
{
    using namespace tmafwd;
    __shared__ __align__(128) float s_filt[STAGES][FILT_FLOATS];
    __shared__ __align__(128) float s_flow[STAGES][FLOW_FLOATS];
    __shared__ __align__(8) uint64_t s_full[STAGES];

    const int C = CT > 0 ? CT : Crt;
    const int tid = threadIdx.x;
    const int tx = tid % TW, ty = tid / TW;
    const int sp = ty * TW + tx;
    const size_t HW = (size_t)H * W;
    const int tiles_per_image = tiles_x * tiles_y;

    if (tid == 0) {
        prefetch_tensormap(&map_flow);
        prefetch_tensormap(&map_filt);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int stage) {   // called by thread 0 only
        const int b = div_tiles_image.quot(tile), rem = tile - b * tiles_per_image;
        const int by = div_tiles_x.quot(rem), bx = rem - by * tiles_x;
        mbar_arrive_expect_tx(&s_full[stage], STAGE_BYTES);
        tma_load_3d(s_flow[stage], &map_flow, &s_full[stage], bx * TW, by * TH, b * 2);
        tma_load_3d(s_filt[stage], &map_filt, &s_full[stage], bx * TW, by * TH, b * 16);
    };

    int tile = blockIdx.x;
    if (tid == 0 && tile < num_tiles) issue(tile, 0);

    for (int it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
        const int stage = it % STAGES;
        const int next = tile + gridDim.x;
        // the buffer of the next stage was last read in iteration it-1, which ended with __syncthreads()
        if (tid == 0 && next < num_tiles) issue(next, (it + 1) % STAGES);

        const int b = div_tiles_image.quot(tile), rem = tile - b * tiles_per_image;
        const int by = div_tiles_x.quot(rem), bx = rem - by * tiles_x;
        const int w_i = bx * TW + tx, h_i = by * TH + ty;
        const size_t pix = (size_t)h_i * W + w_i;
        const float *img = in1 + (size_t)b * C * HW;
        float *o = out + (size_t)b * C * HW + pix;

        mbar_wait(&s_full[stage], (uint32_t)((it / STAGES) & 1));

        if (w_i < W && h_i < H) {
            const float fx = s_flow[stage][sp], fy = s_flow[stage][TH * TW + sp];
            const FiPix p = fi_pixel(w_i, h_i, fx, fy, W, H, 4);
            if (!p.in_range) {   // :2814-2819 copies input1
                for (int c = 0; c < C; ++c) st_stream(o + (size_t)c * HW, __ldg(img + (size_t)c * HW + pix));
            } else {
                float w[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) w[k] = s_filt[stage][k * (TH * TW) + sp];
                const float qTL = (1 - p.alpha) * (1 - p.beta), qTR = p.alpha * (1 - p.beta);
                const float qBL = (1 - p.alpha) * p.beta, qBR = p.alpha * p.beta;
                if (p.L >= 0 && p.T >= 0 && p.L + 3 < W && p.T + 3 < H) {
                    // interior window: no clamping, rows are contiguous -> immediate-offset loads
                    const float *win = img + ((size_t)p.T * W + p.L);
                    if (CT > 0) {
                        float res[CT > 0 ? CT : 1];
#pragma unroll
                        for (int c = 0; c < CT; ++c) res[c] = fi_window_interior(win + (size_t)c * HW, W, w, qTL, qTR, qBL, qBR);
#pragma unroll
                        for (int c = 0; c < CT; ++c) st_stream(o + (size_t)c * HW, res[c]);
                    } else {
#pragma unroll 2
                        for (int c = 0; c < C; ++c)
                            st_stream(o + (size_t)c * HW, fi_window_interior(win + (size_t)c * HW, W, w, qTL, qTR, qBL, qBR));
                    }
                } else {
                    // window touches the border: clamp every tap (:2751-2753)
                    int ro[4], co[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ro[j] = clampi(p.T + j, 0, H - 1) * W;
                        co[j] = clampi(p.L + j, 0, W - 1);
                    }
                    for (int c = 0; c < C; ++c) {
                        const float *pl = img + (size_t)c * HW;
                        float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] += __ldg(pl + ro[j] + co[i]) * w[j * 4 + i];
                        st_stream(o + (size_t)c * HW, qTL * Q[0] + qTR * Q[1] + qBL * Q[2] + qBR * Q[3]);
                    }
                }
            }
        }
        __syncthreads();   // every thread is done with this stage's buffers before they are refilled
    }
}

// ------------------------------------------------------------------------------------------
// DKR forward (4-input, "_deforconv", "_nofilterwithdeforconv").  Taps outermost: the deformed
// sampling geometry of a tap is computed once and applied to CCH channels.  out = sum over taps of
// q(quadrant) * S * w, which is the reference's q_TL*TL + ... regrouped per tap.
// ------------------------------------------------------------------------------------------
template <int V, int FT, int CCH>
__global__ void __launch_bounds__(BX *BY, MINB_FWD_DKR)
fi_forward_dkr_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                      const float *__restrict__ in4, float *__restrict__ out, int C, int H, int W, int Frt)
{
    const int F = FT > 0 ? FT : Frt;
    const int T2 = F * F;
    const int w_i = blockIdx.x * BX + threadIdx.x;
    const int h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)h_i * W + w_i;
    const float *img = in1 + (size_t)b * C * HW;
    float *o = out + (size_t)b * C * HW + pix;

    if (V == V_DKR && !(F == 4 || F == 6)) {  // forward gate of the 4-input family (:68): output stays zero
        for (int c = 0; c < C; ++c) st_stream(o + (size_t)c * HW, 0.0f);
        return;
    }
    const float fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
    const FiPix p = fi_pixel(w_i, h_i, fx, fy, W, H, F);
    if (!p.in_range) {  // :225-230 copies input1
        for (int c = 0; c < C; ++c) st_stream(o + (size_t)c * HW, __ldg(img + (size_t)c * HW + pix));
        return;
    }
    const float *wp = (V == V_NOFILT) ? nullptr : in3 + (size_t)b * T2 * HW + pix;
    const float *op = (V == V_NOFILT ? in3 : in4) + (size_t)b * 2 * T2 * HW + pix;

    for (int c0 = 0; c0 < C; c0 += CCH) {
        float acc[CCH];
#pragma unroll
        for (int cc = 0; cc < CCH; ++cc) acc[cc] = 0.0f;
        const float *pl = img + (size_t)c0 * HW;
#pragma unroll 1
        for (int j = 0; j < F; ++j) {
            const int cy = clampi(p.T + j, 0, H - 1);
#pragma unroll
            for (int i = 0; i < (FT > 0 ? FT : F); ++i) {
                const int cx = clampi(p.L + i, 0, W - 1);
                const int k = j * F + i;
                const float wgt = (V == V_NOFILT) ? 1.0f : __ldg(wp + (size_t)k * HW);
                const float oy = __ldg(op + (size_t)k * HW), ox = __ldg(op + (size_t)(T2 + k) * HW);
                const Deform d = fi_deform(cy, cx, oy, ox, p, H, W);
                const bool top = (V == V_DKR) ? (j < F / 2) : d.top;
                const bool left = (V == V_DKR) ? (i < F / 2) : d.left;
                const QuadCoef qc = quad_coef(top, left, p.alpha, p.beta);
                const float PTL = (1 - d.phiX) * (1 - d.phiY), PTR = d.phiX * (1 - d.phiY);
                const float PBL = (1 - d.phiX) * d.phiY, PBR = d.phiY * d.phiX;
#pragma unroll
                for (int cc = 0; cc < CCH; ++cc)
                    if (c0 + cc < C) {
                        const float *q = pl + (size_t)cc * HW;
                        const float S = PTL * __ldg(q + d.aTL) + PTR * __ldg(q + d.aTR) +
                                        PBL * __ldg(q + d.aBL) + PBR * __ldg(q + d.aBR);   // :110-111
                        acc[cc] += qc.q * (S * wgt);
                    }
            }
        }
#pragma unroll
        for (int cc = 0; cc < CCH; ++cc)
            if (c0 + cc < C) st_stream(o + (size_t)(c0 + cc) * HW, acc[cc]);
    }
}

// ------------------------------------------------------------------------------------------
// backward, all four families: ONE pass over the taps per chunk of CCH channels (the reference makes
// three passes for "_ori" and five for the DKR families).
//   gi1  (image)   : RED scatter to the undeformed clamped tap -- the only real scatter
//   gi2  (flow)    : register accumulation, one store per component
//   gi3  (filter)  : thread-private: stored once per tap (first chunk) / own-pixel += (later chunks)
//   goff (offsets) : likewise; gradinput3 for V_NOFILT, gradinput4 otherwise
// gi1 must be zero on entry (the launcher clears it on the stream); nothing else needs clearing.
// ------------------------------------------------------------------------------------------
template <int V, int FT, int CCH>
__global__ void __launch_bounds__(BX *BY, minb_bwd(V))
fi_backward_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                   const float *__restrict__ in4, const float *__restrict__ gout, float *__restrict__ gi1,
                   float *__restrict__ gi2, float *__restrict__ gi3, float *__restrict__ gi4,
                   int C, int H, int W, int Frt)
{
    const int F = FT > 0 ? FT : Frt;
    const int T2 = F * F;
    const int w_i = blockIdx.x * BX + threadIdx.x;
    const int h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)h_i * W + w_i;

    const float fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix);
    const FiPix p = fi_pixel(w_i, h_i, fx, fy, W, H, F);

    float *g2 = gi2 + (size_t)b * 2 * HW + pix;
    float *g3 = (V == V_NOFILT) ? nullptr : gi3 + (size_t)b * T2 * HW + pix;
    float *go = (V == V_ORI) ? nullptr : (V == V_NOFILT ? gi3 : gi4) + (size_t)b * 2 * T2 * HW + pix;

    if (!p.in_range) {  // contributes nothing; the reference leaves the caller's zeros (:2863)
        st_stream(g2, 0.0f);
        st_stream(g2 + HW, 0.0f);
        for (int k = 0; k < T2; ++k) {
            if (V != V_NOFILT) st_stream(g3 + (size_t)k * HW, 0.0f);
            if (V != V_ORI) { st_stream(go + (size_t)k * HW, 0.0f); st_stream(go + (size_t)(T2 + k) * HW, 0.0f); }
        }
        return;
    }

    const float *img = in1 + (size_t)b * C * HW;
    float *gimg = gi1 + (size_t)b * C * HW;
    const float *gop = gout + (size_t)b * C * HW + pix;
    const float *wp = (V == V_NOFILT) ? nullptr : in3 + (size_t)b * T2 * HW + pix;
    const float *op = (V == V_ORI) ? nullptr
                      : (V == V_NOFILT ? in3 : in4) + (size_t)b * 2 * T2 * HW + pix;
    float gx = 0.0f, gy = 0.0f;

    for (int c0 = 0; c0 < C; c0 += CCH) {
        float g[CCH];
#pragma unroll
        for (int cc = 0; cc < CCH; ++cc) g[cc] = (c0 + cc < C) ? ld_stream(gop + (size_t)(c0 + cc) * HW) : 0.0f;
        const float *pl = img + (size_t)c0 * HW;
        float *gpl = gimg + (size_t)c0 * HW;
        const bool first = (c0 == 0);

#pragma unroll 1
        for (int j = 0; j < F; ++j) {
            const int cy = clampi(p.T + j, 0, H - 1);
            // Everything one tap row reads from HBM (filter taps, offsets) is requested before the first use: issued
            // tap by tap, each load sat behind the previous tap's REDs / stores and a warp paid 16 DRAM round trips
            // in sequence (ncu: 58 % of the stall samples on the first use of the tap weight; 2.34 -> 1.23 ms at
            // 1080p x 8).  Requesting a row further ahead was measured slower (2.14 ms).
            constexpr int FR = FT > 0 ? FT : 1;
            float wrow[FR], oyrow[FR], oxrow[FR];
            if (FT > 0) {
#pragma unroll
                for (int i = 0; i < FR; ++i) {
                    const int k = j * F + i;
                    wrow[i] = (V == V_NOFILT) ? 1.0f : ld_stream(wp + (size_t)k * HW);
                    if (V != V_ORI) { oyrow[i] = ld_stream(op + (size_t)k * HW); oxrow[i] = ld_stream(op + (size_t)(T2 + k) * HW); }
                }
            }
            float vrow[FR][CCH];
            if (FT > 0 && V == V_ORI) {
#pragma unroll
                for (int i = 0; i < FT; ++i) {
                    const int a = cy * W + clampi(p.L + i, 0, W - 1);
#pragma unroll
                    for (int cc = 0; cc < CCH; ++cc) vrow[i][cc] = (c0 + cc < C) ? __ldg(pl + (size_t)cc * HW + a) : 0.0f;
                }
            }
#pragma unroll
            for (int i = 0; i < (FT > 0 ? FT : F); ++i) {
                const int cx = clampi(p.L + i, 0, W - 1);
                const int k = j * F + i;
                const int a = cy * W + cx;
                const float wgt = (V == V_NOFILT) ? 1.0f : (FT > 0 ? wrow[FT > 0 ? i : 0] : __ldg(wp + (size_t)k * HW));
                float s3 = 0.0f, soy = 0.0f, sox = 0.0f;
                if (V == V_ORI) {
                    const QuadCoef qc = quad_coef(j < F / 2, i < F / 2, p.alpha, p.beta);
#pragma unroll
                    for (int cc = 0; cc < CCH; ++cc)
                        if (c0 + cc < C) {
                            const float gq = g[cc] * qc.q;                       // TL_grad (:2885)
                            const float v = FT > 0 ? vrow[FT > 0 ? i : 0][cc] : __ldg(pl + (size_t)cc * HW + a);
                            red_add(gpl + (size_t)cc * HW + a, gq * wgt);       // :2890-2892
                            s3 += gq * v;                                       // :2893-2895
                            const float t = g[cc] * (v * wgt);
                            gx += qc.cx * t;
                            gy += qc.cy * t;
                        }
                } else {
                    const float oy = FT > 0 ? oyrow[FT > 0 ? i : 0] : __ldg(op + (size_t)k * HW);
                    const float ox = FT > 0 ? oxrow[FT > 0 ? i : 0] : __ldg(op + (size_t)(T2 + k) * HW);
                    const Deform d = fi_deform(cy, cx, oy, ox, p, H, W);
                    const bool top = (V == V_DKR) ? (j < F / 2) : d.top;
                    const bool left = (V == V_DKR) ? (i < F / 2) : d.left;
                    const QuadCoef qc = quad_coef(top, left, p.alpha, p.beta);
                    const float PTL = (1 - d.phiX) * (1 - d.phiY), PTR = d.phiX * (1 - d.phiY);
                    const float PBL = (1 - d.phiX) * d.phiY, PBR = d.phiY * d.phiX;
                    float vc[CCH][4];
#pragma unroll
                    for (int cc = 0; cc < CCH; ++cc)
                        if (c0 + cc < C) {
                            const float *qp = pl + (size_t)cc * HW;
                            vc[cc][0] = __ldg(qp + d.aTL); vc[cc][1] = __ldg(qp + d.aTR);
                            vc[cc][2] = __ldg(qp + d.aBL); vc[cc][3] = __ldg(qp + d.aBR);
                        }
#pragma unroll
                    for (int cc = 0; cc < CCH; ++cc)
                        if (c0 + cc < C) {
                            const float vTL = vc[cc][0], vTR = vc[cc][1], vBL = vc[cc][2], vBR = vc[cc][3];
                            const float S = PTL * vTL + PTR * vTR + PBL * vBL + PBR * vBR;
                            const float dSy = -(1 - d.phiX) * vTL + (1 - d.phiX) * vBL - d.phiX * vTR + d.phiX * vBR;  // :986-989
                            const float dSx = -(1 - d.phiY) * vTL + (1 - d.phiY) * vTR - d.phiY * vBL + d.phiY * vBR;  // :1104-1107
                            const float gq = g[cc] * qc.q;
                            red_add(gpl + (size_t)cc * HW + a, gq * wgt);   // undeformed tap (:497-499, :2258)
                            s3 += gq * S;                                   // :520-522
                            soy += gq * dSy * wgt;                          // :990-993
                            sox += gq * dSx * wgt;                          // :1108-1111
                            const float t = g[cc] * (S * wgt);
                            gx += qc.cx * t;
                            gy += qc.cy * t;
                        }
                }
                // thread-private gradients: plain stores, accumulate across channel chunks in place
                if (V != V_NOFILT) {
                    float *q3 = g3 + (size_t)k * HW;
                    *q3 = first ? s3 : *q3 + s3;
                }
                if (V != V_ORI) {
                    float *qy = go + (size_t)k * HW, *qx = go + (size_t)(T2 + k) * HW;
                    *qy = first ? soy : *qy + soy;
                    *qx = first ? sox : *qx + sox;
                }
            }
        }
    }
    st_stream(g2, gx);
    st_stream(g2 + HW, gy);
}

// ---- launchers -----------------------------------------------------------------------------
template <int V>
int launch_forward(const float *in1, const float *in2, const float *in3, const float *in4, float *out,
                   int B, int C, int H, int W, int F, cudaStream_t s, float scale = 1.0f, int accumulate = 0, size_t out_bs = 0)
{
    if (out_bs == 0) out_bs = (size_t)C * H * W;   // dense output
    const bool blend = scale != 1.0f || accumulate != 0 || out_bs != (size_t)C * H * W;   // "_ori" only: strip kernel or direct kernel
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || F <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!in1 || !in2 || !in3 || !out || ((V == V_DKR || V == V_DEFOR) && !in4)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if constexpr (V == V_ORI) {
        if (F == 4) {
            const int path = forced_forward_path();
            if (path == PATH_AUTO && C > 4) {
                // many channels (context features): shared-memory regions streamed channel group by channel group
                const int e = fi_bigc_forward_ori(in1, in2, in3, out, B, C, H, W, scale, accumulate, out_bs, s);
                if (e >= 0) return e;
            }
            if (path == PATH_AUTO || path == PATH_STRIP) {
                // production path: strip-walking kernel, image gathers from a rolling shared-memory window
                const int e = fi_strip_forward_ori(in1, in2, in3, out, B, C, H, W, scale, accumulate, out_bs, s);
                if (e >= 0) return e;
            }
            // TMA-streamed taps, gathers through L1; needs 16-byte aligned rows for the tensor maps
            if (W % 4 == 0 && aligned16(in2) && aligned16(in3) && path != PATH_DIRECT && !blend) {
                using namespace tmafwd;
                CUtensorMap mflow, mfilt;
                if (encode_tensor_map_3d(&mflow, in2, W, H, (uint64_t)B * 2, TW, TH, 2) &&
                    encode_tensor_map_3d(&mfilt, in3, W, H, (uint64_t)B * 16, TW, TH, 16)) {
                    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
                    const long long num_tiles = (long long)tiles_x * tiles_y * B;
                    if (num_tiles < (1ll << 31)) {
                        const int nblk = (int)std::min<long long>(num_tiles, (long long)sm_count() * 4);
                        const FastDiv dx((unsigned)tiles_x), di((unsigned)(tiles_x * tiles_y));
                        if (C == 3)
                            fi_forward_ori_tma_kernel<3><<<nblk, NTHREADS, 0, s>>>(mflow, mfilt, in1, out, C, H, W, tiles_x,
                                                                                  tiles_y, (int)num_tiles, dx, di);
                        else
                            fi_forward_ori_tma_kernel<0><<<nblk, NTHREADS, 0, s>>>(mflow, mfilt, in1, out, C, H, W, tiles_x,
                                                                                  tiles_y, (int)num_tiles, dx, di);
                        note_launch();
                        return check_launch("filterinterpolation forward (tma)");
                    }
                }
            }
            fi_forward_ori_kernel<4><<<grid, block, 0, s>>>(in1, in2, in3, out, C, H, W, F, scale, accumulate, out_bs);
        } else {
            fi_forward_ori_kernel<0><<<grid, block, 0, s>>>(in1, in2, in3, out, C, H, W, F, scale, accumulate, out_bs);
        }
    } else if (F == 4) {
        if (forced_forward_path() != PATH_DIRECT) {
            // production path: strip-walking kernel, bilinear samples from the rolling shared-memory window
            const int e = (V == V_NOFILT) ? fi_strip_forward_dkr(V, in1, in2, nullptr, in3, out, B, C, H, W, s)
                                          : fi_strip_forward_dkr(V, in1, in2, in3, in4, out, B, C, H, W, s);
            if (e >= 0) return e;
        }
        if (C == 3) fi_forward_dkr_kernel<V, 4, 3><<<grid, block, 0, s>>>(in1, in2, in3, in4, out, C, H, W, F);
        else        fi_forward_dkr_kernel<V, 4, 4><<<grid, block, 0, s>>>(in1, in2, in3, in4, out, C, H, W, F);
    } else {
        fi_forward_dkr_kernel<V, 0, 4><<<grid, block, 0, s>>>(in1, in2, in3, in4, out, C, H, W, F);
    }
    note_launch();
    return check_launch("filterinterpolation forward");
}

template <int V>
int launch_backward(const float *in1, const float *in2, const float *in3, const float *in4, const float *gout,
                    float *gi1, float *gi2, float *gi3, float *gi4,
                    int B, int C, int H, int W, int F, cudaStream_t s)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || F <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!in1 || !in2 || !in3 || !gout || !gi1 || !gi2 || !gi3) return VFIDKR_ERR_ARG;
    if ((V == V_DKR || V == V_DEFOR) && (!in4 || !gi4)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    int e = set_error(cudaMemsetAsync(gi1, 0, sizeof(float) * (size_t)B * C * H * W, s), "clear gradinput1");
    if (e) return e;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    if (F == 4) {
        if (C == 3) fi_backward_kernel<V, 4, 3><<<grid, block, 0, s>>>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, C, H, W, F);
        else        fi_backward_kernel<V, 4, 4><<<grid, block, 0, s>>>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, C, H, W, F);
    } else {
        fi_backward_kernel<V, 0, 4><<<grid, block, 0, s>>>(in1, in2, in3, in4, gout, gi1, gi2, gi3, gi4, C, H, W, F);
    }
    note_launch();
    return check_launch("filterinterpolation backward");
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_debug_force_forward_path(int path)
{
    if (path < PATH_AUTO || path > PATH_DIRECT) return -1;
    (void)forced_forward_path();      // consult the environment first, so that it cannot override this call later
    return g_forced_path.exchange(path, std::memory_order_relaxed);
}

VFIDKR_API int vfidkr_filterinterpolation_forward_ori(const float *i1, const float *i2, const float *i3, float *out,
                                                      int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_forward<V_ORI>(i1, i2, i3, nullptr, out, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_forward_ori_blend(const float *i1, const float *i2, const float *i3, float *out,
                                                            int B, int C, int H, int W, int F, float scale, int accumulate,
                                                            long long out_batch_stride, vfidkr_stream_t s)
{
    if (out_batch_stride != 0 && out_batch_stride < (long long)C * H * W) return VFIDKR_ERR_ARG;
    return launch_forward<V_ORI>(i1, i2, i3, nullptr, out, B, C, H, W, F, (cudaStream_t)s, scale, accumulate != 0, (size_t)out_batch_stride);
}

VFIDKR_API int vfidkr_filterinterpolation_backward_ori(const float *i1, const float *i2, const float *i3,
                                                       const float *g, float *gi1, float *gi2, float *gi3,
                                                       int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_backward<V_ORI>(i1, i2, i3, nullptr, g, gi1, gi2, gi3, nullptr, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_forward_dkr(const float *i1, const float *i2, const float *i3,
                                                      const float *i4, float *out,
                                                      int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_forward<V_DKR>(i1, i2, i3, i4, out, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_backward_dkr(const float *i1, const float *i2, const float *i3,
                                                       const float *i4, const float *g, float *gi1, float *gi2,
                                                       float *gi3, float *gi4,
                                                       int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_backward<V_DKR>(i1, i2, i3, i4, g, gi1, gi2, gi3, gi4, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_forward_deforconv(const float *i1, const float *i2, const float *i3,
                                                            const float *i4, float *out,
                                                            int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_forward<V_DEFOR>(i1, i2, i3, i4, out, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_backward_deforconv(const float *i1, const float *i2, const float *i3,
                                                             const float *i4, const float *g, float *gi1,
                                                             float *gi2, float *gi3, float *gi4,
                                                             int B, int C, int H, int W, int F, vfidkr_stream_t s)
{ return launch_backward<V_DEFOR>(i1, i2, i3, i4, g, gi1, gi2, gi3, gi4, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_forward_nofilterwithdeforconv(const float *i1, const float *i2,
                                                                        const float *i3, float *out,
                                                                        int B, int C, int H, int W, int F,
                                                                        vfidkr_stream_t s)
{ return launch_forward<V_NOFILT>(i1, i2, i3, nullptr, out, B, C, H, W, F, (cudaStream_t)s); }

VFIDKR_API int vfidkr_filterinterpolation_backward_nofilterwithdeforconv(const float *i1, const float *i2,
                                                                         const float *i3, const float *g,
                                                                         float *gi1, float *gi2, float *gi3,
                                                                         int B, int C, int H, int W, int F,
                                                                         vfidkr_stream_t s)
{ return launch_backward<V_NOFILT>(i1, i2, i3, nullptr, g, gi1, gi2, gi3, nullptr, B, C, H, W, F, (cudaStream_t)s); }
