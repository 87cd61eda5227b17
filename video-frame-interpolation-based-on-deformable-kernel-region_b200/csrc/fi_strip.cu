// fi_strip.cu -- FilterInterpolation "_ori" forward, strip-walking kernel with a rolling shared-memory image window.
//
// What is computed follows my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu:2692-2823.
// How: the op streams 18 planes per pixel (flow + 16 filter taps) that are read exactly once, and gathers a
// 4 x 4 window of the image around (x + fx, y + fy) per channel.  Gathering 48 scalars per pixel through L1
// is what bounds the straightforward kernels (L1 tag/sector throughput and 1.8 cycles per LDG of LSU issue),
// so here the gathers are served from SHARED memory:
//
//   * one persistent CTA per SM walks TW x 4 pixel tiles (TW = 144, or 128 in the second instantiation) DOWN a strip
//     segment of one frame; segments are drawn by the CTAs in an order that keeps neighbouring CTAs on neighbouring strips
//     at the same rows at the same time (their overlapping window columns then hit in L2 instead of coming from HBM twice);
//   * TW / 8 compute warps (18 / 16), one pixel per thread, threads numbered along the tile's rows.  A thread loads the
//     flow of ITS pixel for the tile LEAD (= 4) tiles ahead straight into registers, derives the clamped extent of that
//     pixel's gather window and folds it into the tile's bounding box (warp min/max reduction + 4 shared-memory atomics
//     per warp);
//   * a producer warp (a) streams the 16 filter planes of the next tiles into a ring of 3 (4) stages with TMA
//     (cp.async.bulk.tensor, ~100 KB in flight per SM = several times HBM latency x bandwidth), and (b) as soon as a
//     tile's bounding box is complete tops up a ROLLING WINDOW of image rows -- a ring of 48 rows x C channels x 192 (160)
//     columns, row y living in slot y % 48, also filled by TMA -- with just the rows the tile adds (about 4 per tile, so
//     the image is fetched from L2 ~1.3x instead of the ~4-6x of per-tile halos), four tiles ahead;
//   * the compute warps then read taps and the 16 x C window values with LDS (conflict-free when the flow is
//     locally smooth) and write the result with streaming stores;
//   * hand-off is mbarrier based (filter full / tile done / bbox done / image full).  The window is re-based
//     (new x origin, refilled) when the flow leaves it, and a tile whose box does not fit at all falls back to
//     clamped global gathers -- correctness never depends on the flow.
//
// Preconditions for this path (checked by the launcher, which otherwise reports "not applicable" and the
// caller uses the generic kernels): F == 4, 1 <= C <= 4, W % 4 == 0, W >= the window width (192 / 160), 16-byte
// aligned bases.
#include <algorithm>
#include <climits>

// Tile width of THIS translation unit.  The file is compiled twice: as itself (144 columns: 18 compute warps + the producer
// = 19 warps, five per scheduler at most, which is what a 96-register budget allows -- 7.6 % faster at 1080p than 128
// columns / 16 warps, whose 17th warp cost the same budget) and through fi_strip_w128.cu (128 columns, window of 160
// columns), which serves images narrower than the 192-column window of this one.  filterinterpolation.cu picks per launch.
#ifndef VFIDKR_ORI_TW
#define VFIDKR_ORI_TW 144
#define VFIDKR_ORI_ENTRY fi_strip_forward_ori_w144
#define VFIDKR_STRIP_NS strip_w144
#define VFIDKR_ORI_PRIMARY 1
#endif
#define VFIDKR_STRIP_TW VFIDKR_ORI_TW
#include "fi_strip_common.cuh"

namespace vfidkr {
namespace VFIDKR_STRIP_NS {

// Optional pipeline statistics (build with -DVFIDKR_STRIP_STATS; read with vfidkr_debug_strip_stats): cycles the
// producer spends in each kind of wait and the compute warps at the two "full" barriers.  Off in production builds.
#ifdef VFIDKR_STRIP_STATS
__device__ unsigned long long g_stats[16];
#define STAT_DECL unsigned long long stat_local[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long stat_begin = clock64()
#define STAT_TIME(slot, stmt) do { const long long t_ = clock64(); stmt; stat_local[slot] += (unsigned long long)(clock64() - t_); } while (0)
#define STAT_INC(slot) (++stat_local[slot])
#define STAT_FLUSH(base, total_slot) do { stat_local[total_slot] = (unsigned long long)(clock64() - stat_begin); \
        for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_stats[(base) + i_], stat_local[i_]); } while (0)
#else
#define STAT_DECL ((void)0)
#define STAT_TIME(slot, stmt) do { stmt; } while (0)
#define STAT_INC(slot) ((void)0)
#define STAT_FLUSH(base, total_slot) ((void)0)
#endif

constexpr int FILT_FLOATS = 16 * NPIX;
constexpr uint32_t FILT_BYTES = FILT_FLOATS * sizeof(float);
// filter pipeline depth: as deep as shared memory allows next to the window ring
// (5 stages with a 35-row window were measured: no gain on smooth flows -- bytes in flight are not the limit -- and
// 8 % slower on the bench flow, whose boxes need the rows)
constexpr size_t SMEM_LIMIT = 227 * 1024;
template <int CG> static __host__ __device__ constexpr int stages()
{
    // as many as fit next to the window ring, at most 4
    int n = (int)((SMEM_LIMIT - 1024 - (size_t)RROWS * row_floats<CG>() * sizeof(float)) / FILT_BYTES);
    return n > 4 ? 4 : n;
}
template <int CG> static __host__ __device__ constexpr size_t smem_bytes()
{
    return (size_t)stages<CG>() * FILT_BYTES + (size_t)RROWS * row_floats<CG>() * sizeof(float) + 1024;
}

// 16 taps x one channel from the rolling window.  off[j] = float offset of (row j of the window, first column);
// INTERIOR: the four columns are consecutive (no border clamp), so every LDS carries an immediate offset.
template <bool INTERIOR>
__device__ __forceinline__ float window_from_smem(const float *__restrict__ ring, const int (&off)[4], const int (&co)[4],
                                                  const float (&w)[16], float qTL, float qTR, float qBL, float qBR)
{
    float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v = INTERIOR ? ring[off[j] + i] : ring[off[j] + co[i]];
            Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] = fmaf(v, w[j * 4 + i], Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)]);
        }
    return qTL * Q[0] + qTR * Q[1] + qBL * Q[2] + qBR * Q[3];
}

template <int CG, bool BLEND>
__global__ void __launch_bounds__(NTHREADS, 1)
fi_forward_ori_strip_kernel(const __grid_constant__ CUtensorMap map_filt, const __grid_constant__ CUtensorMap map_img,
                            const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                            int H, int W, int tiles_x, int tiles_y, int nseg, int segt, int num_items,
                            const FastDiv div_tiles_x, const FastDiv div_nseg, int *__restrict__ work_counter,
                            float scale, int accumulate, size_t out_bs)
{
    // BLEND: the epilogue of SURVEY.md 8f rank 1 (warp both directions and blend): output = scale * result (+ what output
    // held), batch items out_bs elements apart.  The plain instantiation carries none of it: predicated off, those ~35
    // instructions per tile still took issue slots of warps whose own instruction stream is the critical path.
    constexpr int SF = stages<CG>();
    constexpr int ROWF = row_floats<CG>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_filt = reinterpret_cast<float *>(smem_raw);                              // [SF][16][TH][TW]
    float *s_ring = s_filt + SF * FILT_FLOATS;                                         // [RROWS][CG][WB]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + RROWS * ROWF);
    // tile_done is a ring of NB (> LEAD) barriers of its own: the producer waits on the last LEAD tiles, which must map to
    // distinct barriers -- indexed by filter stage it aliased tiles t-4 and t-1 whenever a kernel has fewer than 4 stages
    uint64_t *filt_full = s_bar, *filt_free = s_bar + SF, *tile_done = s_bar + 2 * SF;   // [SF], [SF], [NB]
    uint64_t *bbox_done = tile_done + NB, *img_full = bbox_done + NB;                    // [NB], [NB]
    Box *s_box = reinterpret_cast<Box *>(img_full + NB);                              // [NB]
    TileMeta *s_meta = reinterpret_cast<TileMeta *>(s_box + NB);                      // [NB]
    int *s_need = reinterpret_cast<int *>(s_meta + NB);                               // [NB], producer private
    ItemQueue *s_queue = reinterpret_cast<ItemQueue *>(s_need + NB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
        prefetch_tensormap(&map_filt);
        prefetch_tensormap(&map_img);
        for (int s = 0; s < SF; ++s) {
            mbar_init(&filt_full[s], 1);
            mbar_init(&filt_free[s], NCOMP_WARPS);
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(&tile_done[s], NCOMP_WARPS);
            mbar_init(&bbox_done[s], NCOMP_WARPS);
            mbar_init(&img_full[s], 1);
            s_box[s] = Box{INT_MAX, INT_MIN, INT_MAX, INT_MIN};
        }
        s_queue->fetched = 0;
        s_queue->wanted = 0;
        fence_mbar_init();
    }
    __syncthreads();

    // k-th item of this CTA -> batch item, column block, first tile row; false when the work is exhausted.
    // The CTA's pipeline slots are the tiles of its items back to back; slots past the end of a ragged last segment
    // of a strip are "null" (no loads, no pixels) so that stage and phase accounting stays uniform.
    auto decode_item = [&](int item_no, int &b, int &bx, int &ty0) -> bool {
        // every lane reads the same entry; the redux hands it to the compiler as a warp-UNIFORM value, so the tile
        // bookkeeping derived from it (loop trip counts, item boundaries) stays on the uniform datapath
        const int item = __reduce_max_sync(0xffffffffu, queue_get(s_queue, item_no));
        if (item < 0) return false;
        const int bs = div_tiles_x.quot(item);   // b * nseg + seg
        bx = item - bs * tiles_x;
        b = div_nseg.quot(bs);
        ty0 = (bs - b * nseg) * segt;
        return true;
    };

    if (warp == NCOMP_WARPS) {
        // ================================ producer warp ================================
        // Two independent streams: the filter planes run as far ahead as the SF-deep ring allows (they only wait
        // for "tile done"), the image window follows the bounding boxes.  The filter stream is advanced from
        // inside every wait of the image stream, so a window re-base never stalls the HBM prefetch.  The same pump
        // serves the compute warps' requests for the next work item.
        STAT_DECL;   // producer: 0 box wait, 1 re-base drain, 2 slot-reuse wait, 3 re-bases, 4 global tiles, 5 tiles, 6 total
        bool exhausted = false;
        int item_no = -1;                                                   // image stream: current item
        auto draw_items = [&](int upto) {
            if (lane == 0) queue_fill(s_queue, work_counter, num_items, upto, exhausted);
            __syncwarp();
        };
        int f_t = 0, f_b = 0, f_bx = 0, f_ty = 0, f_left = 0, f_item = -1;   // filter stream position
        bool f_end = false;
        auto pump_filters = [&]() {
            draw_items(*(volatile int *)&s_queue->wanted);
            while (!f_end && (f_t < SF || mbar_test(&filt_free[f_t % SF], (uint32_t)(((f_t / SF) - 1) & 1)))) {
                if (f_left == 0) {
                    ++f_item;
                    draw_items(f_item + 1);
                    if (!decode_item(f_item, f_b, f_bx, f_ty)) { f_end = true; break; }
                    f_left = segt;
                } else {
                    ++f_ty;
                }
                --f_left;
                if (lane == 0) {
                    const int sf = f_t % SF;
                    if (f_ty < tiles_y) {
                        mbar_arrive_expect_tx(&filt_full[sf], FILT_BYTES);
                        tma_load_3d(s_filt + sf * FILT_FLOATS, &map_filt, &filt_full[sf], f_bx * TW, f_ty * TH, f_b * 16);
                    } else {
                        mbar_arrive(&filt_full[sf]);   // null slot
                    }
                }
                ++f_t;
            }
        };
        // wait for a barrier phase while keeping the filter stream going; traps instead of hanging on a protocol error
        auto wait_pumping = [&](uint64_t *bar, uint32_t parity) {
            for (uint32_t spins = 0; !mbar_try_wait_hint(bar, parity, 100u); ++spins) {
                pump_filters();
                if (spins > (1u << 26)) __trap();   // several seconds of polling: a protocol error, not a slow neighbour
            }
        };

        // Window state.  Rows [max(base, hi - RROWS), hi) of the current window are resident; row y lives at UNWRAPPED ring
        // position wr0 + (y - base), i.e. in slot (that position) mod RROWS.  The ring is written strictly in position
        // order, also across window restarts (new x origin, new item, jump): a restart just continues at the write
        // position, and the tiles still in flight keep reading their own window through their own descriptor
        // (x origin + slot offset per tile).  Nothing is drained; a load only waits for the in-flight tiles whose rows
        // it would overwrite.  (Draining at every restart cost ~7.5 us per work item, 14 % of the kernel at zero flow.)
        int xorg = 0, base = 0, hi = 0, wr0 = 0;
        int b = 0, bx = 0, ty = 0, left = 0;
        // The window of one item must never serve the next: `stale` is raised at every item start and cleared only by
        // a re-base.  (It has to survive tiles that do not touch the window -- MODE_GLOBAL / MODE_NONE -- or the first
        // shared-memory tile after them could find the previous item's rows "resident".)
        bool stale = true;
#pragma unroll 1
        for (int t = 0;; ++t) {
            if (left == 0) {
                ++item_no;
                draw_items(item_no + 1);
                if (!decode_item(item_no, b, bx, ty)) break;
                left = segt;
                stale = true;
            } else {
                ++ty;
            }
            --left;
            const int sb = t % NB;
            pump_filters();
            // ---- image window of tile t, as soon as the compute warps have folded its bounding box ----
            STAT_TIME(0, wait_pumping(&bbox_done[sb], (uint32_t)((t / NB) & 1)));
            STAT_INC(5);
            const Box bb = s_box[sb];
            __syncwarp();
            if (lane == 0) s_box[sb] = Box{INT_MAX, INT_MIN, INT_MAX, INT_MIN};   // ready for tile t + NB

            int mode = MODE_NONE, my_need = INT_MAX;
            int load_lo = 0, load_hi = -1;
            if (bb.xmax >= bb.xmin) {
                const int width = bb.xmax - bb.xmin + 1, slack = WB - width;
                if (bb.ymax - bb.ymin + 1 > RROWS || slack < 7) {
                    mode = MODE_GLOBAL;   // the tile's box does not fit the window at all
                    STAT_INC(4);
                } else {
                    mode = MODE_SMEM;
                    const bool rebase = stale || bb.xmin < xorg || bb.xmax >= xorg + WB || bb.ymin < max(base, hi - RROWS);
                    if (rebase || bb.ymin > hi) {
                        // restart the window at this tile (nothing of it is resident yet); the ring continues
                        const int wr = wr0 + (hi - base);
                        if (rebase) {
                            STAT_INC(3);
                            xorg = (bb.xmin - (slack >= 14 ? slack / 2 : 0)) & ~7;   // sector-aligned, box centred when it can be
                            stale = false;
                        }
                        base = hi = bb.ymin;
                        wr0 = wr;
                    }
                    // Writing ring position p reuses the slot of position p - RROWS: every tile still in flight must be past
                    // it.  bbox_done(t) means every compute warp has started tile t - LEAD: older tiles are consumed.  (Only
                    // these last LEAD tiles may be waited on: their tile_done phase is the barrier's current one.)
                    my_need = wr0 + (bb.ymin - base);
                    const int wr_after = wr0 + (max(hi, bb.ymax + 1) - base);
                    int oldest = max(0, t - LEAD);
                    for (;;) {
                        int need = my_need;
                        for (int q = oldest; q < t; ++q) need = min(need, s_need[q % NB]);
                        if (wr_after - need <= RROWS) break;
                        STAT_TIME(2, wait_pumping(&tile_done[oldest % NB], (uint32_t)((oldest / NB) & 1)));   // oldest < t here
                        ++oldest;
                    }
                    load_lo = max(hi, bb.ymin);
                    load_hi = bb.ymax;
                    hi = max(hi, bb.ymax + 1);
                }
            }
            // slot of row y of this tile's window: (y + soff) mod RROWS
            const int soff = (int)((unsigned)(((wr0 - base) % RROWS) + RROWS) % RROWS);
            if (lane == 0) {
                s_need[sb] = my_need;
                s_meta[sb].mode = mode;
                s_meta[sb].xorg = xorg;
                s_meta[sb].soff = soff;
                const int nrows = max(load_hi - load_lo + 1, 0);
                mbar_arrive_expect_tx(&img_full[sb], (uint32_t)nrows * ROWF * (uint32_t)sizeof(float));   // release: publishes the descriptor
                int slot = (int)((unsigned)(load_lo + soff) % RROWS);
                for (int y = load_lo; y <= load_hi; ++y) {
                    tma_load_4d(s_ring + slot * ROWF, &map_img, &img_full[sb], xorg, y, 0, b);
                    slot = slot + 1 == RROWS ? 0 : slot + 1;
                }
            }
            __syncwarp();
        }
        while (!f_end) {   // filter planes of the last tiles whose ring slots were still busy
            if (f_t >= SF) mbar_wait_sleepy(&filt_free[f_t % SF], (uint32_t)(((f_t / SF) - 1) & 1));
            pump_filters();
        }
        if (lane == 0) STAT_FLUSH(0, 6);
    } else {
        // ================================ compute warps ================================
        const int tx = tid % TW, tyy = tid / TW;   // position inside the tile
        const unsigned tile_step = (unsigned)(TH * W);
        // 32-bit shared addresses of the barrier / box rings, computed once
        const uint32_t a_filt_full = smem_u32(filt_full), a_tile_done = smem_u32(tile_done), a_filt_free = smem_u32(filt_free);
        const uint32_t a_bbox_done = smem_u32(bbox_done), a_img_full = smem_u32(img_full), a_box = smem_u32(s_box);

        // left == 0 marks the end of the work: the cursor stays there (no pixel, no further queue reads)
        auto start_item = [&](Cursor &c) {
            int bx, ty0;
            if (!decode_item(c.item_no, c.b, bx, ty0)) { c.left = 0; c.b = 0; c.w_i = 0; c.h_i = INT_MAX / 2; c.pix = 0; return; }
            c.left = segt;
            c.w_i = bx * TW + tx;
            c.h_i = ty0 * TH + tyy;   // rows past the frame (ragged last segment) are >= H by construction
            c.pix = (unsigned)(c.h_i * W + c.w_i);
        };
        // returns true when the cursor entered a new item (its running pointers have to be re-derived)
        auto advance = [&](Cursor &c) -> bool {
            if (c.left == 0) return false;
            if (--c.left == 0) { ++c.item_no; start_item(c); return true; }
            c.h_i += TH; c.pix += tile_step;
            return false;
        };
        auto has_pixel = [&](const Cursor &c) { return c.w_i < W && c.h_i < H; };

        // three cursors over the same sequence: the tile being computed, the tile whose box is folded (LEAD ahead)
        // and the tile whose flow is requested (LEAD + 1 ahead)
        STAT_DECL;   // compute warps: 0 image-full wait, 1 filter-full wait, 2 tiles, 3 total
        Cursor cur{0, 0, 0, 0, 0, 0};
        start_item(cur);
        Cursor fold = cur, req = cur;
        // (Running 64-bit pointers stepped per tile instead of these per-tile products were measured: 25 fewer instructions
        // per tile, 20-28 bytes of spills at the 96-register cap, -1 % on smooth and +3 % on rough flows.  Not kept.)
        const size_t out_item = BLEND ? out_bs : (size_t)CG * HW;
        auto flow_ptr = [&](const Cursor &c) { return in2 + ((size_t)c.b * 2 * HW + c.pix); };
        auto out_ptr = [&](const Cursor &c) { return out + ((size_t)c.b * out_item + c.pix); };
        auto advance_req = [&]() { advance(req); };

        auto request_flow = [&](const Cursor &c, float &fx, float &fy) {
            fx = 0.0f; fy = 0.0f;
            if (has_pixel(c)) {
                const float *f = flow_ptr(c);
                fx = ld_stream(f);
                fy = ld_stream(f + HW);
            }
        };
        // Turns the flow of the cursor's pixel into (x2, y2) -- x2 = -1 marks "out of range" (:2735-2736), which is
        // all the tile computation needs later -- and folds the clamped gather window into the slot's bounding box.
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
            if (c.left == 0) return;   // past the end of the work: no such slot
            xmin = __reduce_min_sync(0xffffffffu, xmin);
            xmax = __reduce_max_sync(0xffffffffu, xmax);
            ymin = __reduce_min_sync(0xffffffffu, ymin);
            ymax = __reduce_max_sync(0xffffffffu, ymax);
            if (lane == 0) {
                const uint32_t sl = (uint32_t)(slot % NB);
                if (xmax >= xmin) {
                    const uint32_t ab = a_box + sl * (uint32_t)sizeof(Box);
                    red_shared_min_a(ab, xmin); red_shared_max_a(ab + 4, xmax);
                    red_shared_min_a(ab + 8, ymin); red_shared_max_a(ab + 12, ymax);
                }
                mbar_arrive_a(a_bbox_done + sl * 8);   // release: the reductions above are visible to the producer
            }
        };

        // prologue: (x2, y2) and boxes of the first LEAD slots, and the flow request for slot LEAD in flight
        float qx[LEAD + 1], qy[LEAD + 1], nx, ny;
#pragma unroll
        for (int k = 0; k < LEAD; ++k) {
            request_flow(req, qx[k], qy[k]);
            advance_req();
        }
        request_flow(req, nx, ny);
        advance_req();
#pragma unroll
        for (int k = 0; k < LEAD; ++k) {
            fold_box(fold, k, qx[k], qy[k]);
            advance(fold);
        }

#pragma unroll 1
        for (int j = 0; cur.left != 0; ++j) {
            // box of slot j + LEAD from the flow requested one tile ago; then request the flow of slot j + LEAD + 1
            qx[LEAD] = nx; qy[LEAD] = ny;
            fold_box(fold, j + LEAD, qx[LEAD], qy[LEAD]);
            advance(fold);
            request_flow(req, nx, ny);
            advance_req();

            const int sf = j % SF, sb = j % NB;
            const float *ft = s_filt + sf * FILT_FLOATS + tid;
            // blend mode: what the output holds is requested now, a whole tile's work before it is added
            float prev[CG];
#pragma unroll
            for (int c = 0; c < CG; ++c) prev[c] = 0.0f;
            if (BLEND && accumulate && has_pixel(cur)) {
                const float *op = out_ptr(cur);
#pragma unroll
                for (int c = 0; c < CG; ++c) prev[c] = __ldcs(op + (size_t)c * HW);
            }

            // Both barriers are probed back to back (a try_wait answers after ~90 cycles even when the phase is long
            // complete: probing them one after the other put two such latencies on every tile of every warp), the taps are
            // requested as soon as the filter stage is there, and only then does the warp wait for its image rows.
            const uint32_t par_i = (uint32_t)((j / NB) & 1), par_f = (uint32_t)((j / SF) & 1);
            const bool ok_f = mbar_try_wait_a(a_filt_full + sf * 8, par_f);
            const bool ok_i = mbar_try_wait_a(a_img_full + sb * 8, par_i);
            if (!ok_f) STAT_TIME(1, mbar_wait_backoff_a(a_filt_full + sf * 8, par_f));
            // The 16 taps go to registers at once and the stage is handed back BEFORE the window arithmetic: the kernel
            // is bound by the bytes it keeps in flight (4 stages x 32 KB per SM; a stage that waits for the slowest
            // warp to finish the whole tile spends a third of its life neither in flight nor in use).
            float w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = ft[k * NPIX];
            if (!ok_i) STAT_TIME(0, mbar_wait_backoff_a(a_img_full + sb * 8, par_i));
            STAT_INC(2);
            const int mode = s_meta[sb].mode, xorg = s_meta[sb].xorg, soff = s_meta[sb].soff;
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_filt_free + sf * 8);   // release: this warp's reads of stage sf are complete

            if (has_pixel(cur)) {
                float *o = out_ptr(cur);
                const float x2 = qx[0], y2 = qy[0];
                if (x2 < 0.0f) {   // out of range: :2814-2819 copies input1
                    const float *img = in1 + (size_t)cur.b * CG * HW + cur.pix;
#pragma unroll
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, BLEND ? scale * __ldg(img + (size_t)c * HW) + prev[c] : __ldg(img + (size_t)c * HW));
                } else {
                    const int ix = (int)x2, iy = (int)y2;
                    const int L = ix - 1, T = iy - 1;                       // window origin for F = 4 (:2745-2748)
                    const float alpha = __fsub_rn(x2, (float)ix), beta = __fsub_rn(y2, (float)iy);
                    const float qTL = (1 - alpha) * (1 - beta), qTR = alpha * (1 - beta);
                    const float qBL = (1 - alpha) * beta, qBR = alpha * beta;
                    float res[CG];
                    if (mode == MODE_SMEM) {
                        int off[4], co[4];
                        if (L >= 0 && T >= 0 && L + 3 < W && T + 3 < H) {
                            // interior window: rows T..T+3 are consecutive ring slots (mod RROWS), columns are contiguous
                            const int r0 = (int)((unsigned)(T + soff) % RROWS);
                            const int o0 = r0 * ROWF + (L - xorg);
#pragma unroll
                            for (int r = 0; r < 4; ++r) off[r] = o0 + r * ROWF - (r0 + r >= RROWS ? RROWS * ROWF : 0);
                            co[0] = co[1] = co[2] = co[3] = 0;
#pragma unroll
                            for (int c = 0; c < CG; ++c)
                                res[c] = window_from_smem<true>(s_ring + c * WB, off, co, w, qTL, qTR, qBL, qBR);
#ifdef VFIDKR_BOUNDS_CHECK
                            for (int c = 0; c < CG; ++c)
                                for (int r = 0; r < 4; ++r)
                                    for (int i = 0; i < 4; ++i) {
                                        const int idx = c * WB + off[r] + i;
                                        const bool inside = idx >= 0 && idx < RROWS * ROWF;
                                        const float want = __ldg(in1 + (size_t)cur.b * CG * HW + (size_t)c * HW + (size_t)(T + r) * W + (L + i));
                                        bounds_check(inside && __float_as_uint(s_ring[inside ? idx : 0]) == __float_as_uint(want));
                                    }
#endif
                        } else {
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                off[r] = (int)((unsigned)(clampi(T + r, 0, H - 1) + soff) % RROWS) * ROWF - xorg;   // :2751
                                co[r] = clampi(L + r, 0, W - 1);                                           // :2753
                            }
#pragma unroll
                            for (int c = 0; c < CG; ++c)
                                res[c] = window_from_smem<false>(s_ring + c * WB, off, co, w, qTL, qTR, qBL, qBR);
#ifdef VFIDKR_BOUNDS_CHECK
                            for (int c = 0; c < CG; ++c)
                                for (int r = 0; r < 4; ++r)
                                    for (int i = 0; i < 4; ++i) {
                                        const int idx = c * WB + off[r] + co[i];
                                        const bool inside = idx >= 0 && idx < RROWS * ROWF;
                                        const float want = __ldg(in1 + (size_t)cur.b * CG * HW + (size_t)c * HW +
                                                                 (size_t)clampi(T + r, 0, H - 1) * W + clampi(L + i, 0, W - 1));
                                        bounds_check(inside && __float_as_uint(s_ring[inside ? idx : 0]) == __float_as_uint(want));
                                    }
#endif
                        }
                    } else {
                        // the tile's windows do not fit the rolling window: clamped gathers from global memory
                        const float *img = in1 + (size_t)cur.b * CG * HW;
                        int ro[4], co[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            ro[r] = clampi(T + r, 0, H - 1) * W;
                            co[r] = clampi(L + r, 0, W - 1);
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
                    for (int c = 0; c < CG; ++c) st_stream(o + (size_t)c * HW, BLEND ? scale * res[c] + prev[c] : res[c]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_a(a_tile_done + sb * 8);   // this warp is done with the window rows of tile j
            advance(cur);
#pragma unroll
            for (int k = 0; k < LEAD; ++k) { qx[k] = qx[k + 1]; qy[k] = qy[k + 1]; }
        }
        if (tid == 0) STAT_FLUSH(8, 3);
    }
}

template <int CG, bool BLEND>
static int launch(const CUtensorMap &mfilt, const float *in1, const float *in2, float *out, int B, int H, int W,
                  float scale, int accumulate, size_t out_bs, cudaStream_t s)
{
    CUtensorMap mimg;
    if (!encode_tensor_map_4d(&mimg, in1, W, H, CG, B, WB, 1, CG)) return -1;
    if ((long long)CG * H * W >= (1ll << 31)) return -1;   // 32-bit pixel offsets inside one batch item
    const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
    if ((long long)tiles_x * tiles_y * B >= (1ll << 28)) return -1;
    const int sms = sm_count();
    const int nseg = choose_segments(B, tiles_x, tiles_y, sms), segt = (tiles_y + nseg - 1) / nseg;
    const long long items = (long long)B * tiles_x * nseg;
    auto kernel = fi_forward_ori_strip_kernel<CG, BLEND>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<CG>()) != cudaSuccess) {
        (void)cudaGetLastError();   // smaller shared-memory carve-out (MIG, ...): let the caller run the generic kernels
        return -1;
    }
    const int nblk = (int)std::min<long long>(sms, items);
    void *counter = nullptr;   // the global work counter: 4 bytes of stream-ordered scratch, zeroed on the stream
    int e = stream_scratch_alloc(&counter, sizeof(int), s);
    if (e) return e;
    e = set_error(cudaMemsetAsync(counter, 0, sizeof(int), s), "clear work counter");
    if (!e) {
        kernel<<<nblk, NTHREADS, smem_bytes<CG>(), s>>>(mfilt, mimg, in1, in2, out, H, W, tiles_x, tiles_y, nseg, segt, (int)items,
                                                       FastDiv((unsigned)tiles_x), FastDiv((unsigned)nseg), static_cast<int *>(counter),
                                                       scale, accumulate, out_bs);
        note_launch();
        e = check_launch("filterinterpolation forward (strip)");
    }
    const int e2 = set_error(cudaFreeAsync(counter, s), "free work counter");
    return e ? e : e2;
}

}  // namespace VFIDKR_STRIP_NS

#if defined(VFIDKR_STRIP_STATS) && defined(VFIDKR_ORI_PRIMARY)
// debug builds only: reads and clears the pipeline statistics (16 counters, see STAT_DECL comments)
extern "C" __attribute__((visibility("default"))) int vfidkr_debug_strip_stats(unsigned long long *out16)
{
    unsigned long long zero[16] = {0};
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(out16, VFIDKR_STRIP_NS::g_stats, sizeof zero) != cudaSuccess) return 1;
    return cudaMemcpyToSymbol(VFIDKR_STRIP_NS::g_stats, zero, sizeof zero) != cudaSuccess;
}
#endif

// Returns VFIDKR_OK / VFIDKR_ERR_CUDA when the strip kernel was launched, -1 when it does not apply.
int VFIDKR_ORI_ENTRY(const float *in1, const float *in2, const float *in3, float *out,
                         int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s)
{
    using namespace VFIDKR_STRIP_NS;
    if (C < 1 || C > 4 || W % 4 != 0 || W < WB) return -1;
    if (!aligned16(in1) || !aligned16(in3)) return -1;
    CUtensorMap mfilt;
    if (!encode_tensor_map_3d(&mfilt, in3, W, H, (uint64_t)B * 16, TW, TH, 16)) return -1;
    const bool blend = scale != 1.0f || accumulate != 0 || out_bs != (size_t)C * H * W;
    if (blend) {
        switch (C) {
        case 1: return launch<1, true>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
        case 2: return launch<2, true>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
        case 3: return launch<3, true>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
        default: return launch<4, true>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
        }
    }
    switch (C) {
    case 1: return launch<1, false>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
    case 2: return launch<2, false>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
    case 3: return launch<3, false>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
    default: return launch<4, false>(mfilt, in1, in2, out, B, H, W, scale, accumulate, out_bs, s);
    }
}

}  // namespace vfidkr

#ifdef VFIDKR_ORI_PRIMARY
VFIDKR_BOUNDS_ACCESSOR(bounds_counts_strip_w144)
#else
VFIDKR_BOUNDS_ACCESSOR(bounds_counts_strip_w128)
#endif
