// fi_strip_common.cuh -- tile geometry and small types shared by the strip-walking FilterInterpolation kernels
// (fi_strip.cu: "_ori" forward; fi_strip_dkr.cu: the deformable-kernel-region forwards).  See fi_strip.cu for the
// description of the scheme (persistent CTA per SM, rolling shared-memory image window, mbarrier pipeline).
#pragma once

#include <climits>

#include "common.cuh"
#include "fi_common.cuh"
#include "tma.cuh"

namespace vfidkr {
namespace strip {

constexpr int TW = 128, TH = 4, NPIX = TW * TH;    // tile = 512 pixels, one per compute thread
constexpr int NCOMP_WARPS = NPIX / 32;             // 16 compute warps
constexpr int NTHREADS = NPIX + 32;                // + 1 producer warp
constexpr int LEAD = 3;                            // flow / bounding box / image window run this many tiles ahead
constexpr int NB = 8;                              // ring of bounding boxes and tile descriptors (> LEAD)
constexpr int WB = 160;                            // columns held by the rolling window (tile + 16 either side)
constexpr int RROWS = 48;                          // rows held by the rolling window (a ring indexed by y % RROWS)
enum { MODE_NONE = 0, MODE_SMEM = 1, MODE_GLOBAL = 2 };

template <int CG> __host__ __device__ constexpr int row_floats() { return CG * WB; }   // one window row: [C][WB]

struct TileMeta { int mode, xorg; };
struct Box { int xmin, xmax, ymin, ymax; };

// Position of one thread in the CTA's sequence of pipeline slots (tiles of its items, back to back), advanced
// incrementally: the per-item decode (two divisions) runs once per item, not once per tile.
struct Cursor {
    int item_no;    // index into this CTA's items
    int left;       // tiles left in the current item, including the current one
    int b;          // batch item
    int w_i, h_i;   // pixel of this thread; h_i >= H marks "no pixel" (null slot / past the end)
    unsigned pix;   // h_i * W + w_i
};

// Split every strip into `nseg` segments so that the items (b, seg, bx) fill whole rounds of one CTA per SM.
// Cost model: rounds x (tiles per segment + ~3 tiles' worth of window refill at each segment start).
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

}  // namespace strip
}  // namespace vfidkr
