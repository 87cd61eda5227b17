// fi_strip_dkr.cu -- forward of the three deformable-kernel-region (DKR) FilterInterpolation families on the
// strip-walking / rolling-window scheme of fi_strip.cu.
//
// What is computed follows my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu
//   :29-426 (4-input, static quadrants)  :1353-1496 ("_deforconv", data-dependent quadrants)
//   :2070-2191 ("_nofilterwithdeforconv": unit weights, offsets in input3).
// Per output pixel the op streams 16 filter taps + 32 tap offsets (224 B/px with the flow and 3 channels) and takes,
// for each of the 16 taps, a BILINEAR sample of every channel at (clamped tap position + learned offset): 192
// gathers per pixel at C = 3.  Served from L1 those gathers bound the direct kernel (2.7 ms at 1080p x 8); here
//   * the image lives in the same rolling shared-memory window as in fi_strip.cu (48 rows x C x 160 columns, TMA
//     row boxes, re-based when the flow leaves it), widened by one row / column either side for the bilinear reach;
//   * the 48 streamed planes go through shared memory in TAP-ROW GROUPS: sub-stage g of a tile holds the 4 taps of
//     window row g -- 4 filter planes, 4 y-offset planes, 4 x-offset planes, 24 KB -- in a 5-deep TMA ring, so the
//     pipeline is 5 x 24 KB deep instead of 1 x 96 KB;
//   * for an in-contract tap (the deformed position stays within one pixel of the tap, i.e. Top in {cy-1, cy},
//     Left in {cx-1, cx}) the two ring rows are picked from three precomputed row offsets with a select -- no
//     modulo per tap -- and the 4 x C values come from the window with LDS; any other tap (wild offsets) reads its
//     four corners from global memory with the same clamp-to-plane rule as the direct kernel, so the result never
//     depends on the offsets being small.
// Preconditions as for fi_strip.cu (F == 4, C <= 4, W % 4 == 0, W >= 160, aligned bases); otherwise the caller
// falls back to the direct kernels of filterinterpolation.cu.
#include <algorithm>

// tile width of THESE kernels (fi_strip.cu chooses its own)
#ifndef VFIDKR_DKR_TW
#define VFIDKR_DKR_TW 128
#endif
#define VFIDKR_STRIP_TW VFIDKR_DKR_TW
#define VFIDKR_STRIP_NS strip_dkr
#ifndef VFIDKR_DKR_MAX_SUB
#define VFIDKR_DKR_MAX_SUB 5
#endif
#include "fi_strip_common.cuh"

namespace vfidkr {
namespace VFIDKR_STRIP_NS {

constexpr int LEAD_D = 3;                                        // lookahead of THIS kernel: 4 (fi_strip.cu's) costs registers it does not have
constexpr int SUB_PLANES = 12;                                   // 4 filter + 4 offY + 4 offX planes per sub-stage
constexpr int SUB_FLOATS = SUB_PLANES * NPIX;
constexpr int GROUP_FLOATS = 4 * NPIX;                           // one TMA box: 4 planes x tile
constexpr uint32_t GROUP_BYTES = GROUP_FLOATS * sizeof(float);
template <int CG> static __host__ __device__ constexpr int sub_stages()
{
    // as many as fit next to the window ring (227 KB per CTA), at most VFIDKR_DKR_MAX_SUB
    int n = (int)((227 * 1024 - 1024 - (size_t)RROWS * row_floats<CG>() * sizeof(float)) / (SUB_FLOATS * sizeof(float)));
    return n > VFIDKR_DKR_MAX_SUB ? VFIDKR_DKR_MAX_SUB : n;
}
template <int CG> static __host__ __device__ constexpr size_t dkr_smem_bytes()
{
    return (size_t)sub_stages<CG>() * SUB_FLOATS * sizeof(float) + (size_t)RROWS * row_floats<CG>() * sizeof(float) + 1024;
}

template <int V, int CG>
__global__ void __launch_bounds__(NTHREADS, 1)
fi_forward_dkr_strip_kernel(const __grid_constant__ CUtensorMap map_filt, const __grid_constant__ CUtensorMap map_off,
                            const __grid_constant__ CUtensorMap map_img,
                            const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                            int H, int W, int tiles_x, int tiles_y, int nseg, int segt, int num_items,
                            const FastDiv div_tiles_x, const FastDiv div_nseg)
{
    constexpr int NS = sub_stages<CG>();
    constexpr int ROWF = row_floats<CG>();
    constexpr bool HAS_FILTER = V != V_NOFILT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_sub = reinterpret_cast<float *>(smem_raw);                               // [NS][12][TH][TW]
    float *s_ring = s_sub + NS * SUB_FLOATS;                                           // [RROWS][CG][WB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + RROWS * ROWF);
    uint64_t *sub_full = s_bar, *sub_empty = s_bar + NS, *tile_done = s_bar + 2 * NS;
    uint64_t *bbox_done = tile_done + NB, *img_full = bbox_done + NB;
    Box *s_box = reinterpret_cast<Box *>(img_full + NB);                              // [NB]
    TileMeta *s_meta = reinterpret_cast<TileMeta *>(s_box + NB);                      // [NB]
    int *s_ymin = reinterpret_cast<int *>(s_meta + NB);                               // [NB], producer private

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;

    // work decomposition: identical to fi_strip.cu (strip segments dealt round-robin, column block fastest)
    const int my_items = (int)blockIdx.x < num_items ? (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n = my_items * segt;

    if (tid == 0) {
        prefetch_tensormap(&map_filt);
        prefetch_tensormap(&map_off);
        prefetch_tensormap(&map_img);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&sub_full[s], 1);
            mbar_init(&sub_empty[s], NCOMP_WARPS);
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(&tile_done[s], NCOMP_WARPS);
            mbar_init(&bbox_done[s], NCOMP_WARPS);
            mbar_init(&img_full[s], 1);
            s_box[s] = Box{INT_MAX, INT_MIN, INT_MAX, INT_MIN};
        }
        fence_mbar_init();
    }
    __syncthreads();

    auto decode_item = [&](int item_no, int &b, int &bx, int &ty0) {
        const int item = (int)blockIdx.x + item_no * (int)gridDim.x;
        const int bs = div_tiles_x.quot(item);   // b * nseg + seg
        bx = item - bs * tiles_x;
        b = div_nseg.quot(bs);
        ty0 = (bs - b * nseg) * segt;
    };

    if (warp == NCOMP_WARPS) {
        // ================================ producer warp ================================
        // stream 1: sub-stages (tile t, tap row g) = u = 4 t + g, as far ahead as the NS-deep ring allows
        int s_u = 0, s_b = 0, s_bx = 0, s_ty = 0, s_left = 0, s_item = -1;
        const int total_sub = 4 * n;
        auto pump_substages = [&]() {
            while (s_u < total_sub && (s_u < NS || mbar_test(&sub_empty[s_u % NS], (uint32_t)(((s_u / NS) - 1) & 1)))) {
                const int g = s_u & 3;
                if (g == 0) {
                    if (s_left == 0) { ++s_item; decode_item(s_item, s_b, s_bx, s_ty); s_left = segt; }
                    else ++s_ty;
                    --s_left;
                }
                if (lane == 0) {
                    const int st = s_u % NS;
                    float *dst = s_sub + st * SUB_FLOATS;
                    if (s_ty < tiles_y) {
                        mbar_arrive_expect_tx(&sub_full[st], (HAS_FILTER ? 3u : 2u) * GROUP_BYTES);
                        if (HAS_FILTER)
                            tma_load_3d(dst, &map_filt, &sub_full[st], s_bx * TW, s_ty * TH, s_b * 16 + 4 * g);
                        tma_load_3d(dst + GROUP_FLOATS, &map_off, &sub_full[st], s_bx * TW, s_ty * TH, s_b * 32 + 4 * g);
                        tma_load_3d(dst + 2 * GROUP_FLOATS, &map_off, &sub_full[st], s_bx * TW, s_ty * TH, s_b * 32 + 16 + 4 * g);
                    } else {
                        mbar_arrive(&sub_full[st]);   // null slot
                    }
                }
                ++s_u;
            }
        };
        auto wait_pumping = [&](uint64_t *bar, uint32_t parity) {
            for (uint32_t spins = 0; !mbar_try_wait_hint(bar, parity, 400u); ++spins) {
                pump_substages();
                if (spins > (1u << 24)) __trap();
            }
        };

        // stream 2: the image window, driven by the bounding boxes (see fi_strip.cu)
        int xorg = 0, base = 0, hi = 0;
        int b = 0, bx = 0, ty = 0, left = 0, item_no = -1;
        bool stale = true;   // raised at every item start, cleared only by a re-base (see fi_strip.cu)
        for (int t = 0; t < n; ++t) {
            if (left == 0) { ++item_no; decode_item(item_no, b, bx, ty); left = segt; stale = true; }
            else ++ty;
            --left;
            const int sb = t % NB;
            pump_substages();
            wait_pumping(&bbox_done[sb], (uint32_t)((t / NB) & 1));
            Box bb = s_box[sb];
            __syncwarp();
            if (lane == 0) s_box[sb] = Box{INT_MAX, INT_MIN, INT_MAX, INT_MIN};

            int mode = MODE_NONE, my_ymin = INT_MAX;
            int load_lo = 0, load_hi = -1;
            if (bb.xmax >= bb.xmin) {
                // bilinear reach of an in-contract deformed tap: one row / column either side, clamped to the plane
                bb.xmin = max(bb.xmin - 1, 0); bb.xmax = min(bb.xmax + 1, W - 1);
                bb.ymin = max(bb.ymin - 1, 0); bb.ymax = min(bb.ymax + 1, H - 1);
                const int width = bb.xmax - bb.xmin + 1, slack = WB - width;
                if (bb.ymax - bb.ymin + 1 > RROWS || slack < 7) {
                    mode = MODE_GLOBAL;
                } else {
                    mode = MODE_SMEM;
                    const bool rebase = stale || bb.xmin < xorg || bb.xmax >= xorg + WB || bb.ymin < max(base, hi - RROWS);
                    int oldest = max(0, t - LEAD_D);   // bbox_done(t): every compute warp has started tile t - LEAD_D
                    if (rebase) {
                        for (; oldest < t; ++oldest) wait_pumping(&tile_done[oldest % NB], (uint32_t)((oldest / NB) & 1));
                        xorg = (bb.xmin - (slack >= 14 ? slack / 2 : 0)) & ~7;
                        base = hi = bb.ymin;
                        stale = false;
                    }
                    if (bb.ymin > hi) base = hi = bb.ymin;
                    for (;;) {
                        int need_lo = bb.ymin;
                        for (int q = oldest; q < t; ++q) need_lo = min(need_lo, s_ymin[q % NB]);
                        if (bb.ymax - need_lo + 1 <= RROWS) break;
                        wait_pumping(&tile_done[oldest % NB], (uint32_t)((oldest / NB) & 1));
                        ++oldest;
                    }
                    load_lo = max(hi, bb.ymin);
                    load_hi = bb.ymax;
                    hi = max(hi, bb.ymax + 1);
                    my_ymin = bb.ymin;
                }
            }
            if (lane == 0) {
                s_ymin[sb] = my_ymin;
                s_meta[sb].mode = mode;
                s_meta[sb].xorg = xorg;
                const int nrows = max(load_hi - load_lo + 1, 0);
                mbar_arrive_expect_tx(&img_full[sb], (uint32_t)nrows * ROWF * (uint32_t)sizeof(float));
                int slot = load_lo % RROWS;
                for (int y = load_lo; y <= load_hi; ++y) {
                    tma_load_4d(s_ring + slot * ROWF, &map_img, &img_full[sb], xorg, y, 0, b);
                    slot = slot + 1 == RROWS ? 0 : slot + 1;
                }
            }
            __syncwarp();
        }
        while (s_u < total_sub) {   // sub-stages of the last tiles whose ring slots were still busy
            if (s_u >= NS) mbar_wait_sleepy(&sub_empty[s_u % NS], (uint32_t)(((s_u / NS) - 1) & 1));
            pump_substages();
        }
    } else {
        // ================================ compute warps ================================
        const int tx = tid % TW, tyy = tid / TW;
        const unsigned tile_step = (unsigned)(TH * W);

        auto start_item = [&](Cursor &c) {
            if (c.item_no >= my_items) { c.left = INT_MAX; c.b = 0; c.w_i = 0; c.h_i = INT_MAX / 2; c.pix = 0; return; }
            int bx, ty0;
            decode_item(c.item_no, c.b, bx, ty0);
            c.left = segt;
            c.w_i = bx * TW + tx;
            c.h_i = ty0 * TH + tyy;
            c.pix = (unsigned)(c.h_i * W + c.w_i);
        };
        auto advance = [&](Cursor &c) {
            if (--c.left == 0) { ++c.item_no; start_item(c); }
            else { c.h_i += TH; c.pix += tile_step; }
        };
        auto has_pixel = [&](const Cursor &c) { return c.w_i < W && c.h_i < H; };

        Cursor cur{0, 0, 0, 0, 0, 0};
        start_item(cur);
        Cursor fold = cur, req = cur;

        auto request_flow = [&](const Cursor &c, float &fx, float &fy) {
            fx = 0.0f; fy = 0.0f;
            if (has_pixel(c)) {
                const float *f = in2 + (size_t)c.b * 2 * HW + c.pix;
                fx = ld_stream(f);
                fy = ld_stream(f + HW);
            }
        };
        auto fold_box = [&](const Cursor &c, int slot, float &fx_x2, float &fy_y2) {
            int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
            float x2 = -1.0f, y2 = 0.0f;
            if (has_pixel(c)) {
                const FiPix p = fi_pixel(c.w_i, c.h_i, fx_x2, fy_y2, W, H, 4);
                if (p.in_range) {
                    x2 = p.x2; y2 = p.y2;
                    xmin = max(p.L, 0); xmax = min(p.L + 3, W - 1);
                    ymin = max(p.T, 0); ymax = min(p.T + 3, H - 1);
                }
            }
            fx_x2 = x2; fy_y2 = y2;
            if (slot >= n) return;
            xmin = __reduce_min_sync(0xffffffffu, xmin);
            xmax = __reduce_max_sync(0xffffffffu, xmax);
            ymin = __reduce_min_sync(0xffffffffu, ymin);
            ymax = __reduce_max_sync(0xffffffffu, ymax);
            if (lane == 0) {
                Box *bx_ = &s_box[slot % NB];
                if (xmax >= xmin) {
                    atomicMin(&bx_->xmin, xmin); atomicMax(&bx_->xmax, xmax);
                    atomicMin(&bx_->ymin, ymin); atomicMax(&bx_->ymax, ymax);
                }
                mbar_arrive(&bbox_done[slot % NB]);
            }
        };

        float qx[LEAD_D + 1], qy[LEAD_D + 1], nx, ny;
#pragma unroll
        for (int k = 0; k < LEAD_D; ++k) {
            request_flow(req, qx[k], qy[k]);
            advance(req);
        }
        request_flow(req, nx, ny);
        advance(req);
#pragma unroll
        for (int k = 0; k < LEAD_D; ++k) {
            fold_box(fold, k, qx[k], qy[k]);
            advance(fold);
        }

        for (int j = 0; j < n; ++j) {
            qx[LEAD_D] = nx; qy[LEAD_D] = ny;
            fold_box(fold, j + LEAD_D, qx[LEAD_D], qy[LEAD_D]);
            advance(fold);
            request_flow(req, nx, ny);
            advance(req);

            const int sb = j % NB;
            mbar_wait_sleepy(&img_full[sb], (uint32_t)((j / NB) & 1));
            const int mode = s_meta[sb].mode, xorg = s_meta[sb].xorg;

            const bool pixel = has_pixel(cur);
            const float x2 = qx[0], y2 = qy[0];
            const bool in_range = pixel && x2 >= 0.0f;
            const float *img = in1 + (size_t)cur.b * CG * HW;

            // per-pixel geometry shared by the 16 taps
            const int ix = (int)x2, iy = (int)y2;
            const int L = ix - 1, T = iy - 1;
            const float alpha = __fsub_rn(x2, (float)ix), beta = __fsub_rn(y2, (float)iy);
            const float qTL = (1 - alpha) * (1 - beta), qTR = alpha * (1 - beta);
            const float qBL = (1 - alpha) * beta, qBR = alpha * beta;
            int cxi[4];          // clamped tap columns (:101,:1393)
            float cxf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { cxi[i] = clampi(L + i, 0, W - 1); cxf[i] = (float)cxi[i]; }
            float acc[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) acc[c] = 0.0f;

#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                const int u = 4 * j + g, st = u % NS;
                mbar_wait_sleepy(&sub_full[st], (uint32_t)((u / NS) & 1));
                // geometry of the four taps of this window row: everything the sub-stage holds goes to registers here, and
                // the sub-stage is handed back BEFORE the window arithmetic (cf. the filter-free barrier of fi_strip.cu)
                const int cy = clampi(T + g, 0, H - 1);          // clamped tap row (:100,:1392)
                float qw[4], phiX[4], phiY[4];
                int Top[4], Left[4];
                bool near_all = mode == MODE_SMEM;
                if (in_range) {
                    const float *sp = s_sub + st * SUB_FLOATS + tid;
                    const float cyf = (float)cy;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float wgt = HAS_FILTER ? sp[i * NPIX] : 1.0f;
                        const float oy = sp[(4 + i) * NPIX], ox = sp[(8 + i) * NPIX];
                        const float fracY = __fadd_rn(cyf, oy), fracX = __fadd_rn(cxf[i], ox);   // :98-99
                        Top[i] = (int)fracY; Left[i] = (int)fracX;                               // :102-103
                        phiY[i] = __fsub_rn(fracY, (float)Top[i]); phiX[i] = __fsub_rn(fracX, (float)Left[i]);
                        bool top, left;
                        if (V == V_DKR) { top = g < 2; left = i < 2; }                            // static quadrants
                        else { top = fracY <= y2; left = fracX <= x2; }                           // :1442-1468
                        qw[i] = (top ? (left ? qTL : qTR) : (left ? qBL : qBR)) * wgt;
                        // in-contract tap: Top in {cy-1, cy}, Left in {cx-1, cx} -> all four corners are in the window
                        near_all = near_all && (unsigned)(Top[i] - cy + 1) < 2u && (unsigned)(Left[i] - cxi[i] + 1) < 2u;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sub_empty[st]);   // release: this warp's reads of the sub-stage are complete
                if (in_range) {
                    // ring offsets of rows cy-1, cy, cy+1 (clamped to the plane) -- only used on the window path
                    const int r_m = (int)((unsigned)max(cy - 1, 0) % RROWS) * ROWF - xorg;
                    const int r_0 = (int)((unsigned)cy % RROWS) * ROWF - xorg;
                    const int r_p = (int)((unsigned)min(cy + 1, H - 1) % RROWS) * ROWF - xorg;
                    if (near_all) {
                        // branch-free block: 16 x C window loads in flight together
                        float v[4][4][CG];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // Top >= 0 by truncation, Bottom = min(Top+1, H-1); Left = -1 only for a wild offset at column 0
                            const int rt = Top[i] < cy ? r_m : r_0, rbm = Top[i] < cy ? r_0 : r_p;
                            const int cl = max(Left[i], 0), cr = min(Left[i] + 1, W - 1);
                            const float *pTL = s_ring + rt + cl, *pTR = s_ring + rt + cr;
                            const float *pBL = s_ring + rbm + cl, *pBR = s_ring + rbm + cr;
#pragma unroll
                            for (int c = 0; c < CG; ++c) {
                                v[i][0][c] = pTL[c * WB]; v[i][1][c] = pTR[c * WB];
                                v[i][2][c] = pBL[c * WB]; v[i][3][c] = pBR[c * WB];
                            }
#ifdef VFIDKR_BOUNDS_CHECK
                            {   // the same four corners by the direct kernel's clamp-to-plane rule, from global memory
                                const int t = clampi(Top[i], 0, H - 1), bm = clampi(min(Top[i], H - 1) + 1, 0, H - 1);
                                const int l = clampi(Left[i], 0, W - 1), r = clampi(min(Left[i], W - 1) + 1, 0, W - 1);
                                const int ids[4] = {rt + cl, rt + cr, rbm + cl, rbm + cr};
                                const int gs[4] = {t * W + l, t * W + r, bm * W + l, bm * W + r};
                                for (int c = 0; c < CG; ++c)
                                    for (int k = 0; k < 4; ++k) {
                                        const int idx = ids[k] + c * WB;
                                        const bool inside = idx >= 0 && idx < RROWS * ROWF;
                                        const float want = __ldg(img + (size_t)c * HW + gs[k]);
                                        bounds_check(inside && __float_as_uint(s_ring[inside ? idx : 0]) == __float_as_uint(want));
                                    }
                            }
#endif
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float PTL = (1 - phiX[i]) * (1 - phiY[i]), PTR = phiX[i] * (1 - phiY[i]);
                            const float PBL = (1 - phiX[i]) * phiY[i], PBR = phiY[i] * phiX[i];
#pragma unroll
                            for (int c = 0; c < CG; ++c) {
                                const float S = PTL * v[i][0][c] + PTR * v[i][1][c] + PBL * v[i][2][c] + PBR * v[i][3][c];   // :110-111
                                acc[c] = fmaf(qw[i], S, acc[c]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float PTL = (1 - phiX[i]) * (1 - phiY[i]), PTR = phiX[i] * (1 - phiY[i]);
                            const float PBL = (1 - phiX[i]) * phiY[i], PBR = phiY[i] * phiX[i];
                            const bool near_tap = mode == MODE_SMEM && (unsigned)(Top[i] - cy + 1) < 2u &&
                                                  (unsigned)(Left[i] - cxi[i] + 1) < 2u;
                            if (near_tap) {
                                const int rt = Top[i] < cy ? r_m : r_0, rbm = Top[i] < cy ? r_0 : r_p;
                                const int cl = max(Left[i], 0), cr = min(Left[i] + 1, W - 1);
#pragma unroll
                                for (int c = 0; c < CG; ++c) {
                                    const float S = PTL * s_ring[rt + cl + c * WB] + PTR * s_ring[rt + cr + c * WB] +
                                                    PBL * s_ring[rbm + cl + c * WB] + PBR * s_ring[rbm + cr + c * WB];
                                    acc[c] = fmaf(qw[i], S, acc[c]);
                                }
                            } else {
                                // wild offset (or a tile without a window): clamp-to-plane corners from global memory
                                const int t = clampi(Top[i], 0, H - 1), bm = clampi(min(Top[i], H - 1) + 1, 0, H - 1);
                                const int l = clampi(Left[i], 0, W - 1), r = clampi(min(Left[i], W - 1) + 1, 0, W - 1);
#pragma unroll
                                for (int c = 0; c < CG; ++c) {
                                    const float *pl = img + (size_t)c * HW;
                                    const float S = PTL * __ldg(pl + t * W + l) + PTR * __ldg(pl + t * W + r) +
                                                    PBL * __ldg(pl + bm * W + l) + PBR * __ldg(pl + bm * W + r);
                                    acc[c] = fmaf(qw[i], S, acc[c]);
                                }
                            }
                        }
                    }
                }
            }
            if (pixel) {
                float *o = out + (size_t)cur.b * CG * HW + cur.pix;
                if (in_range) {
#pragma unroll
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, acc[c]);
                } else {   // :225-230 copies input1
#pragma unroll
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, __ldg(img + cur.pix + (size_t)c * HW));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_done[sb]);
            advance(cur);
#pragma unroll
            for (int k = 0; k < LEAD_D; ++k) { qx[k] = qx[k + 1]; qy[k] = qy[k + 1]; }
        }
    }
}

template <int V, int CG>
static int launch_dkr(const float *in1, const float *in2, const float *filt, const float *offs, float *out,
                      int B, int H, int W, cudaStream_t s)
{
    CUtensorMap mfilt, moff, mimg;
    // V_NOFILT has no filter tensor: the map is unused, encode it over the offsets to keep the argument valid
    if (!encode_tensor_map_3d(&mfilt, V == V_NOFILT ? offs : filt, W, H, (uint64_t)B * (V == V_NOFILT ? 32 : 16), TW, TH, 4)) return -1;
    if (!encode_tensor_map_3d(&moff, offs, W, H, (uint64_t)B * 32, TW, TH, 4)) return -1;
    if (!encode_tensor_map_4d(&mimg, in1, W, H, CG, B, WB, 1, CG)) return -1;
    if ((long long)CG * H * W >= (1ll << 31)) return -1;
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    if ((long long)tiles_x * tiles_y * B >= (1ll << 26)) return -1;
    const int sms = sm_count();
    const int nseg = choose_segments(B, tiles_x, tiles_y, sms), segt = (tiles_y + nseg - 1) / nseg;
    const long long items = (long long)B * tiles_x * nseg;
    auto kernel = fi_forward_dkr_strip_kernel<V, CG>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkr_smem_bytes<CG>()) != cudaSuccess) {
        (void)cudaGetLastError();   // let the caller run the per-pixel kernels
        return -1;
    }
    const int nblk = (int)std::min<long long>(sms, items);
    kernel<<<nblk, NTHREADS, dkr_smem_bytes<CG>(), s>>>(mfilt, moff, mimg, in1, in2, out, H, W, tiles_x, tiles_y, nseg, segt,
                                                       (int)items, FastDiv((unsigned)tiles_x), FastDiv((unsigned)nseg));
    note_launch();
    return check_launch("filterinterpolation DKR forward (strip)");
}

template <int V>
static int dispatch_dkr(const float *in1, const float *in2, const float *filt, const float *offs, float *out,
                        int B, int C, int H, int W, cudaStream_t s)
{
    switch (C) {
    case 1: return launch_dkr<V, 1>(in1, in2, filt, offs, out, B, H, W, s);
    case 2: return launch_dkr<V, 2>(in1, in2, filt, offs, out, B, H, W, s);
    case 3: return launch_dkr<V, 3>(in1, in2, filt, offs, out, B, H, W, s);
    default: return launch_dkr<V, 4>(in1, in2, filt, offs, out, B, H, W, s);
    }
}

}  // namespace VFIDKR_STRIP_NS

// Returns VFIDKR_OK / VFIDKR_ERR_CUDA when the strip kernel was launched, -1 when it does not apply.
// variant: V_DKR / V_DEFOR (filt = input3 [B,16,H,W], offs = input4 [B,32,H,W]) or V_NOFILT (offs = input3, filt unused).
int fi_strip_forward_dkr(int variant, const float *in1, const float *in2, const float *filt, const float *offs, float *out,
                         int B, int C, int H, int W, cudaStream_t s)
{
    using namespace VFIDKR_STRIP_NS;
    if (C < 1 || C > 4 || W % 4 != 0 || W < WB) return -1;
    if (!aligned16(in1) || !aligned16(offs) || (variant != V_NOFILT && !aligned16(filt))) return -1;
    switch (variant) {
    case V_DKR: return dispatch_dkr<V_DKR>(in1, in2, filt, offs, out, B, C, H, W, s);
    case V_DEFOR: return dispatch_dkr<V_DEFOR>(in1, in2, filt, offs, out, B, C, H, W, s);
    case V_NOFILT: return dispatch_dkr<V_NOFILT>(in1, in2, filt, offs, out, B, C, H, W, s);
    default: return -1;
    }
}

}  // namespace vfidkr

VFIDKR_BOUNDS_ACCESSOR(bounds_counts_strip_dkr)
