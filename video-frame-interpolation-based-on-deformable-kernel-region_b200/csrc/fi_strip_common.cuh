// fi_strip_common.cuh -- tile geometry and small types shared by the strip-walking FilterInterpolation kernels
// (fi_strip.cu: "_ori" forward; fi_strip_dkr.cu: the deformable-kernel-region forwards).  See fi_strip.cu for the
// description of the scheme (persistent CTA per SM, rolling shared-memory image window, mbarrier pipeline).
#pragma once

#include <climits>

#include "common.cuh"
#include "fi_common.cuh"
#include "tma.cuh"

// Every translation unit that includes this header names its own namespace (VFIDKR_STRIP_NS): the kernels are templates with
// external linkage, and two units with different tile widths must not share symbols.
#ifndef VFIDKR_STRIP_NS
#define VFIDKR_STRIP_NS strip
#endif

namespace vfidkr {
namespace VFIDKR_STRIP_NS {

// Tile width: a translation unit may choose its own (VFIDKR_STRIP_TW before the include; everything here has internal
// linkage, so the "_ori" and the DKR kernels can differ).  Threads are numbered along the tile's rows; when the width is
// not a multiple of 32 a warp straddles two rows, which nothing below depends on.
#ifndef VFIDKR_STRIP_TW
#define VFIDKR_STRIP_TW 128
#endif
constexpr int TW = VFIDKR_STRIP_TW, TH = 4, NPIX = TW * TH;   // tile: one pixel per compute thread
static_assert(NPIX % 32 == 0 && TW % 8 == 0, "tile = whole warps, TMA boxes of whole sectors");
constexpr int NCOMP_WARPS = NPIX / 32;             // compute warps
constexpr int NTHREADS = NPIX + 32;                // + 1 producer warp
#ifndef VFIDKR_STRIP_LEAD
#define VFIDKR_STRIP_LEAD 4
#endif
constexpr int LEAD = VFIDKR_STRIP_LEAD;                            // flow / bounding box / image window run this many tiles ahead
constexpr int NB = 8;                              // ring of bounding boxes and tile descriptors (> LEAD)
constexpr int WB = (TW + 32 + 31) / 32 * 32;       // columns held by the rolling window: tile + >= 16 either side; a multiple of 32 so
                                                   // that every row slot of the ring ([C][WB] floats) is a 128-byte aligned TMA destination
#ifndef VFIDKR_STRIP_RROWS
#define VFIDKR_STRIP_RROWS 48
#endif
constexpr int RROWS = VFIDKR_STRIP_RROWS;           // rows held by the rolling window (a ring indexed by y % RROWS)
enum { MODE_NONE = 0, MODE_SMEM = 1, MODE_GLOBAL = 2 };

template <int CG> static __host__ __device__ constexpr int row_floats() { return CG * WB; }   // one window row: [C][WB]

struct TileMeta { int mode, xorg, soff, pad_; };   // window descriptor of one tile: x origin and slot offset (row y -> slot (y + soff) mod RROWS)
struct Box { int xmin, xmax, ymin, ymax; };

// Position of one thread in the CTA's sequence of pipeline slots (tiles of its items, back to back), advanced
// incrementally: the per-item decode (two divisions) runs once per item, not once per tile.
struct Cursor {
    int item_no;    // index into this CTA's item sequence
    int left;       // tiles left in the current item, including the current one (0 once the work is exhausted)
    int b;          // batch item
    int w_i, h_i;   // pixel of this thread; h_i >= H marks "no pixel" (null slot / past the end)
    unsigned pix;   // h_i * W + w_i
};

// ---- dynamic work distribution ----------------------------------------------------------------------------------
// Work items are strip segments (b, seg, bx), numbered with bx fastest.  CTAs draw them from a global counter, so a
// CTA that meets expensive tiles (re-bases, bank conflicts, fallbacks) simply draws fewer: measured with static
// round-robin dealing the slowest SM ran 1.15x - 1.5x longer than the average one.  Drawing in item order keeps
// neighbouring SMs on neighbouring strips of the same rows (their window columns overlap and hit in L2).
// Inside a CTA several parties walk the SAME item sequence at different distances ahead (filter stream, image
// stream, three cursors per compute thread), so the drawn ids go through a small shared-memory queue: entry k is
// the k-th item of this CTA, -1 = no more work.  Only lane 0 of the producer warp draws; everybody else reads.
// Items are drawn LAZILY -- entry k only when some party first asks for it (a few tiles before the CTA finishes item
// k-1) -- because every item drawn early is an item another SM cannot take.
constexpr int ITEM_QUEUE = 32;      // ring of item ids (>> the distance between the most and least advanced party)

struct ItemQueue {
    int ids[ITEM_QUEUE];
    int fetched;                    // number of entries published so far
    int wanted;                     // highest entry count any compute thread is waiting for
};

// k-th item of this CTA; -1 when the work is exhausted.  If it has not been drawn yet the caller posts its wish and
// waits for the producer warp (which polls `wanted` from its pump loop, at most ~0.4 us apart).
// (noinline: it runs once per item and cursor; inlined it bloats and fragments the hot tile loop -- measured 11 %)
static __device__ __noinline__ int queue_get(ItemQueue *q, int k)
{
    volatile int *nf = &q->fetched;
    if (*nf <= k) {
        atomicMax(&q->wanted, k + 1);
        while (*nf <= k) __nanosleep(64);
    }
    __threadfence_block();
    return *(volatile int *)&q->ids[k % ITEM_QUEUE];
}
// producer lane 0: draw until entries [0, upto) exist or the work is exhausted
__device__ __forceinline__ void queue_fill(ItemQueue *q, int *counter, int num_items, int upto, bool &exhausted)
{
    int nf = q->fetched;
    while (nf < upto && !exhausted) {
        int id = atomicAdd(counter, 1);
        if (id >= num_items) { id = -1; exhausted = true; }
        q->ids[nf % ITEM_QUEUE] = id;
        __threadfence_block();
        ++nf;
        *(volatile int *)&q->fetched = nf;
    }
}

// Segments per strip.  With uniform tiles the makespan is rounds x item length, rounds = ceil(items / SMs), so the
// split is chosen to fill whole rounds (e.g. 128 strips x 8 segments = 1024 items = 6.9 -> 7 rounds on 148 SMs)
// while keeping segments long: each segment start costs ~3 tiles' worth of window refill.  Dynamic drawing then
// absorbs the non-uniformity of real tiles inside that schedule.
inline int choose_segments(int B, int tiles_x, int tiles_y, int sms)
{
    int best_nseg = 1;
    long long best_cost = LLONG_MAX;
    for (int ns = 1; ns <= (tiles_y < 64 ? tiles_y : 64); ++ns) {
        const long long items = (long long)B * tiles_x * ns;
        const long long rounds = (items + sms - 1) / sms;
        const long long cost = rounds * ((tiles_y + ns - 1) / ns + 3);
        if (cost < best_cost) { best_cost = cost; best_nseg = ns; }
    }
    return best_nseg;
}

}  // namespace VFIDKR_STRIP_NS
}  // namespace vfidkr
