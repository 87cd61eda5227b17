// fi_strip.cu -- FilterInterpolation "_ori" forward, strip-walking kernel with a rolling shared-memory image window.
//
// What is computed follows my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu:2692-2823.
// How: the op streams 18 planes per pixel (flow + 16 filter taps) that are read exactly once, and gathers a
// 4 x 4 window of the image around (x + fx, y + fy) per channel.  Gathering 48 scalars per pixel through L1
// is what bounds the straightforward kernels (L1 tag/sector throughput and 1.8 cycles per LDG of LSU issue),
// so here the gathers are served from SHARED memory:
//
//   * a persistent CTA owns a contiguous run of 64 x 4 pixel tiles, walked DOWN a 64-column strip of one frame;
//   * a producer warp streams the flow + filter planes of the next tiles into a 3-stage ring with TMA
//     (cp.async.bulk.tensor), computes the bounding box of the tile's clamped gather windows from the flow
//     as soon as it lands, and tops up a ROLLING WINDOW of image rows -- a ring of 12 chunks x 4 rows x 96
//     columns x C channels, also filled by TMA -- with just the rows the next tile adds (about 4 per tile, so
//     the image is fetched from L2 ~1.5x instead of the ~4-6x of per-tile halos);
//   * 8 compute warps (one pixel per thread, warps along x) read flow, taps and the 16 x C window values
//     with conflict-free LDS and write the result with streaming stores;
//   * producer/consumer hand-off is mbarrier based (full/empty per stage; image loads complete_tx on their own
//     barrier); the window is re-based (new x origin, refilled) when the flow leaves it, and a tile whose box
//     does not fit at all falls back to clamped global gathers -- correctness never depends on the flow.
//
// Preconditions for this path (checked by the launcher, which otherwise reports "not applicable" and the
// caller uses the generic kernels): F == 4, 1 <= C <= 4, W % 4 == 0, W >= 96, 16-byte aligned bases.
#include <algorithm>
#include <climits>

#include "common.cuh"
#include "fi_common.cuh"
#include "tma.cuh"

namespace vfidkr {
namespace strip {

constexpr int TW = 64, TH = 4, NPIX = TW * TH;     // tile = 256 pixels, one per compute thread
constexpr int NCOMP_WARPS = NPIX / 32;             // 8 compute warps
constexpr int NTHREADS = NPIX + 32;                // + 1 producer warp
constexpr int S = 3;                               // flow+filter pipeline depth
constexpr int WB = 96;                             // columns held by the rolling window
constexpr int RCH = 12;                            // window ring: 12 chunks x 4 rows = 48 rows
constexpr int CROWS = 4;
constexpr int FF_PLANES = 18;                      // 2 flow + 16 taps
constexpr int FF_FLOATS = FF_PLANES * NPIX;
constexpr uint32_t FF_BYTES = FF_FLOATS * sizeof(float);
enum { MODE_NONE = 0, MODE_SMEM = 1, MODE_GLOBAL = 2 };

template <int CG> __host__ __device__ constexpr int chunk_floats() { return CG * CROWS * WB; }
template <int CG> __host__ __device__ constexpr size_t smem_bytes()
{
    return (size_t)S * FF_BYTES + (size_t)RCH * chunk_floats<CG>() * sizeof(float) + 256;
}

struct TileMeta { int mode, xorg; };

// window values of one channel plane from the rolling window; rb[j] = float offset of the (clamped) row j
template <bool INTERIOR_X>
__device__ __forceinline__ float window_from_smem(const float *__restrict__ ring, const int (&rb)[4], const int (&co)[4],
                                                  const float (&w)[16], float qTL, float qTR, float qBL, float qBR)
{
    float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = INTERIOR_X ? ring[rb[j] + co[0] + i] : ring[rb[j] + co[i]];
            Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] = fmaf(v, w[j * 4 + i], Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)]);
        }
    return qTL * Q[0] + qTR * Q[1] + qBL * Q[2] + qBR * Q[3];
}

template <int CG>
__global__ void __launch_bounds__(NTHREADS, 2)
fi_forward_ori_strip_kernel(const __grid_constant__ CUtensorMap map_flow, const __grid_constant__ CUtensorMap map_filt,
                            const __grid_constant__ CUtensorMap map_img,
                            const float *__restrict__ in1, float *__restrict__ out,
                            int H, int W, int tiles_x, int tiles_y, int total_tiles,
                            const FastDiv div_tiles_y, const FastDiv div_tiles_x)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_ff = reinterpret_cast<float *>(smem_raw);                               // [S][18][TH][TW]
    float *s_ring = s_ff + S * FF_FLOATS;                                             // [RCH][CG][4][WB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + RCH * chunk_floats<CG>());
    uint64_t *ff_full = s_bar, *ff_empty = s_bar + S, *img_full = s_bar + 2 * S;
    TileMeta *s_meta = reinterpret_cast<TileMeta *>(s_bar + 3 * S);                   // [S]
    int *s_cmin = reinterpret_cast<int *>(s_meta + S);                                // [S], producer private

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;

    // contiguous, balanced run of tiles for this CTA (tile index = (b * tiles_x + bx) * tiles_y + ty)
    const long long t_begin = (long long)total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = (long long)total_tiles * (blockIdx.x + 1) / gridDim.x;
    const int n = (int)(t_end - t_begin);

    if (tid == 0) {
        prefetch_tensormap(&map_flow);
        prefetch_tensormap(&map_filt);
        prefetch_tensormap(&map_img);
        for (int s = 0; s < S; ++s) {
            mbar_init(&ff_full[s], 1);
            mbar_init(&ff_empty[s], NCOMP_WARPS);
            mbar_init(&img_full[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    auto decode = [&](int t, int &b, int &bx, int &ty, int &strip_id) {
        strip_id = div_tiles_y.quot(t);
        ty = t - strip_id * tiles_y;
        b = div_tiles_x.quot(strip_id);
        bx = strip_id - b * tiles_x;
    };

    if (warp == NCOMP_WARPS) {
        // ================================ producer warp ================================
        int cur_strip = -1, xorg = 0, base = 0, hi = 0;   // window state: chunks [max(base, hi - RCH), hi) are resident
        for (int i = 0; i <= n; ++i) {
            if (i < n) {
                const int stage = i % S;
                if (i >= S) mbar_wait_guarded(&ff_empty[stage], (uint32_t)(((i / S) - 1) & 1));
                if (lane == 0) {
                    int b, bx, ty, sid;
                    decode((int)t_begin + i, b, bx, ty, sid);
                    float *dst = s_ff + stage * FF_FLOATS;
                    mbar_arrive_expect_tx(&ff_full[stage], FF_BYTES);
                    tma_load_3d(dst, &map_flow, &ff_full[stage], bx * TW, ty * TH, b * 2);
                    tma_load_3d(dst + 2 * NPIX, &map_filt, &ff_full[stage], bx * TW, ty * TH, b * 16);
                }
            }
            if (i == 0) continue;
            // ---- image stage of tile j = i - 1 (its flow was requested one iteration ago) ----
            const int j = i - 1, sj = j % S;
            int b, bx, ty, sid;
            decode((int)t_begin + j, b, bx, ty, sid);
            mbar_wait_guarded(&ff_full[sj], (uint32_t)((j / S) & 1));
            const float *fl = s_ff + sj * FF_FLOATS;
            int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
            for (int k = 0; k < NPIX / 32; ++k) {
                const int sp = lane + 32 * k;
                const int w_i = bx * TW + (sp % TW), h_i = ty * TH + (sp / TW);
                if (w_i < W && h_i < H) {
                    const FiPix p = fi_pixel(w_i, h_i, fl[sp], fl[NPIX + sp], W, H, 4);
                    if (p.in_range) {
                        xmin = min(xmin, max(p.L, 0));
                        xmax = max(xmax, min(p.L + 3, W - 1));
                        ymin = min(ymin, max(p.T, 0));
                        ymax = max(ymax, min(p.T + 3, H - 1));
                    }
                }
            }
            xmin = __reduce_min_sync(0xffffffffu, xmin);
            xmax = __reduce_max_sync(0xffffffffu, xmax);
            ymin = __reduce_min_sync(0xffffffffu, ymin);
            ymax = __reduce_max_sync(0xffffffffu, ymax);

            int mode = MODE_NONE, my_cmin = INT_MAX;
            uint32_t bytes = 0;
            int load_lo = 0, load_hi = -1;
            if (xmax >= xmin) {
                const int cmin = ymin >> 2, cmax = ymax >> 2;
                const int width = xmax - xmin + 1, slack = WB - width;
                if (cmax - cmin + 1 > RCH || slack < 7) {
                    mode = MODE_GLOBAL;   // the tile's box does not fit the window at all
                } else {
                    mode = MODE_SMEM;
                    const bool rebase = sid != cur_strip || xmin < xorg || xmax >= xorg + WB || cmin < max(base, hi - RCH);
                    int oldest = max(0, j - S + 1);   // tiles before this one are known to be consumed
                    if (rebase) {
                        // everything in flight may still read the window: drain, then restart it around this tile
                        for (; oldest < j; ++oldest) mbar_wait_guarded(&ff_empty[oldest % S], (uint32_t)((oldest / S) & 1));
                        cur_strip = sid;
                        xorg = (xmin - (slack >= 14 ? slack / 2 : 0)) & ~7;   // sector-aligned, box centred when it can be
                        base = hi = cmin;
                    }
                    if (cmin > hi) base = hi = cmin;   // jumped ahead: nothing older is needed any more
                    // loading chunk c reuses the slot of chunk c - RCH: every tile still in flight must be past it
                    for (;;) {
                        int need_lo = cmin;
                        for (int q = oldest; q < j; ++q) need_lo = min(need_lo, s_cmin[q % S]);
                        if (cmax - need_lo + 1 <= RCH) break;
                        mbar_wait_guarded(&ff_empty[oldest % S], (uint32_t)((oldest / S) & 1));   // oldest < j here
                        ++oldest;
                    }
                    load_lo = max(hi, cmin);
                    load_hi = cmax;
                    if (load_hi >= load_lo) bytes = (uint32_t)(load_hi - load_lo + 1) * chunk_floats<CG>() * sizeof(float);
                    hi = max(hi, cmax + 1);
                    my_cmin = cmin;
                }
            }
            if (lane == 0) {
                s_cmin[sj] = my_cmin;
                s_meta[sj].mode = mode;
                s_meta[sj].xorg = xorg;
                mbar_arrive_expect_tx(&img_full[sj], bytes);   // release: publishes the meta words as well
                for (int c = load_lo; c <= load_hi; ++c)
                    tma_load_4d(s_ring + (c % RCH) * chunk_floats<CG>(), &map_img, &img_full[sj], xorg, c * CROWS, 0, b);
            }
            __syncwarp();
        }
    } else {
        // ================================ compute warps ================================
        const int sp = tid;                       // position inside the tile
        const int tx = sp % TW, tyy = sp / TW;
        for (int j = 0; j < n; ++j) {
            const int sj = j % S;
            const uint32_t ph = (uint32_t)((j / S) & 1);
            int b, bx, ty, sid;
            decode((int)t_begin + j, b, bx, ty, sid);
            const int w_i = bx * TW + tx, h_i = ty * TH + tyy;
            const size_t pix = (size_t)h_i * W + w_i;
            const float *img = in1 + (size_t)b * CG * HW;
            float *o = out + (size_t)b * CG * HW + pix;
            const float *ff = s_ff + sj * FF_FLOATS;

            mbar_wait_guarded(&ff_full[sj], ph);
            mbar_wait_guarded(&img_full[sj], ph);
            const int mode = s_meta[sj].mode, xorg = s_meta[sj].xorg;

            if (w_i < W && h_i < H) {
                const FiPix p = fi_pixel(w_i, h_i, ff[sp], ff[NPIX + sp], W, H, 4);
                if (!p.in_range) {   // :2814-2819 copies input1
#pragma unroll
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, __ldg(img + (size_t)c * HW + pix));
                } else {
                    float w[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) w[k] = ff[(2 + k) * NPIX + sp];
                    const float qTL = (1 - p.alpha) * (1 - p.beta), qTR = p.alpha * (1 - p.beta);
                    const float qBL = (1 - p.alpha) * p.beta, qBR = p.alpha * p.beta;
                    float res[CG];
                    if (mode == MODE_SMEM) {
                        int rb[4], co[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int yy = clampi(p.T + r, 0, H - 1);   // :2751
                            rb[r] = (int)((unsigned)(yy >> 2) % RCH) * chunk_floats<CG>() + (yy & 3) * WB;
                        }
                        if (p.L >= 0 && p.L + 3 < W) {
                            co[0] = p.L - xorg; co[1] = co[2] = co[3] = 0;
#pragma unroll
                            for (int c = 0; c < CG; ++c)
                                res[c] = window_from_smem<true>(s_ring + c * (CROWS * WB), rb, co, w, qTL, qTR, qBL, qBR);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) co[i] = clampi(p.L + i, 0, W - 1) - xorg;   // :2753
#pragma unroll
                            for (int c = 0; c < CG; ++c)
                                res[c] = window_from_smem<false>(s_ring + c * (CROWS * WB), rb, co, w, qTL, qTR, qBL, qBR);
                        }
                    } else {
                        // the tile's windows do not fit the rolling window: clamped gathers from global memory
                        int ro[4], co[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            ro[r] = clampi(p.T + r, 0, H - 1) * W;
                            co[r] = clampi(p.L + r, 0, W - 1);
                        }
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            const float *pl = img + (size_t)c * HW;
                            float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int r = 0; r < 4; ++r)
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    Q[(r < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] += __ldg(pl + ro[r] + co[i]) * w[r * 4 + i];
                            res[c] = qTL * Q[0] + qTR * Q[1] + qBL * Q[2] + qBR * Q[3];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, res[c]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ff_empty[sj]);   // this warp is done with stage sj and with the window rows of tile j
        }
    }
}

template <int CG>
static int launch(const CUtensorMap &mflow, const CUtensorMap &mfilt, const float *in1, float *out, int B, int H, int W,
                  cudaStream_t s)
{
    CUtensorMap mimg;
    if (!encode_tensor_map_4d(&mimg, in1, W, H, CG, B, WB, CROWS, CG)) return -1;
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    const long long total = (long long)tiles_x * tiles_y * B;
    if (total >= (1ll << 31)) return -1;
    auto kernel = fi_forward_ori_strip_kernel<CG>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<CG>());
    // two CTAs per SM; never more CTAs than tiles, and keep runs long enough that the window pays off
    const int nblk = (int)std::max<long long>(1, std::min<long long>((long long)sm_count() * 2, total / 8 + 1));
    kernel<<<nblk, NTHREADS, smem_bytes<CG>(), s>>>(mflow, mfilt, mimg, in1, out, H, W, tiles_x, tiles_y, (int)total,
                                                   FastDiv((unsigned)tiles_y), FastDiv((unsigned)tiles_x));
    note_launch();
    return check_launch("filterinterpolation forward (strip)");
}

}  // namespace strip

// Returns VFIDKR_OK / VFIDKR_ERR_CUDA when the strip kernel was launched, -1 when it does not apply.
int fi_strip_forward_ori(const float *in1, const float *in2, const float *in3, float *out,
                         int B, int C, int H, int W, cudaStream_t s)
{
    using namespace strip;
    if (C < 1 || C > 4 || W % 4 != 0 || W < WB || H < CROWS) return -1;
    if (!aligned16(in1) || !aligned16(in2) || !aligned16(in3)) return -1;
    CUtensorMap mflow, mfilt;
    if (!encode_tensor_map_3d(&mflow, in2, W, H, (uint64_t)B * 2, TW, TH, 2) ||
        !encode_tensor_map_3d(&mfilt, in3, W, H, (uint64_t)B * 16, TW, TH, 16))
        return -1;
    switch (C) {
    case 1: return launch<1>(mflow, mfilt, in1, out, B, H, W, s);
    case 2: return launch<2>(mflow, mfilt, in1, out, B, H, W, s);
    case 3: return launch<3>(mflow, mfilt, in1, out, B, H, W, s);
    default: return launch<4>(mflow, mfilt, in1, out, B, H, W, s);
    }
}

}  // namespace vfidkr
