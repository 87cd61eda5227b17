// projection.cu -- FlowProjection / DepthFlowProjection (forward splat with atomics, averaging,
// hole filling; gather backward) for sm_100a.
//
// Behaviour follows my_package/FlowProjection/flowprojection_cuda_kernel.cu:29-301 and
// my_package/DepthFlowProjection/depthflowprojection_cuda_kernel.cu:29-341.  One templated kernel
// family serves both (DEPTH = false -> weight 1).  Atomics are kept exactly where the reference
// splats (the four corners of (x+fx, y+fy), three planes each); what changes:
//   * every source pixel adds the SAME triple (-d*fx, -d*fy, d) to the 2 x 2 block {T, Bm} x {L, R}, so the splat
//     is split: ONE 128-bit vector RED per pixel (REDG.E.ADD.F32x4) deposits the triple at the block's top-left
//     cell of an interleaved scratch image, and a dense, atomic-free 2 x 2 box pass (fused with the averaging)
//     spreads it -- same sums, 1 atomic instead of 12, which is what bounds this op (the L2 retires roughly
//     one RED request per clock per slice, whatever its width);
//   * the scratch image (16 B per pixel) is stream-ordered memory from the device's default pool
//     (library-private stream-ordered pool, capi.cu: no synchronisation, cached after the first call);
//   * the accumulation planes are cleared by the library on the stream (no caller zero-fill);
//   * the backward is a pure gather with register accumulation and a single store per output
//     (the reference does eight / sixteen read-modify-writes of its own pixel).
#include <algorithm>
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace vfidkr {
namespace {

constexpr int BX = 32, BY = 8;

struct Corners {
    bool in_range;
    int L, T, R, Bm;
};

// flowprojection_cuda_kernel.cu:63-73 -- note: no |flow| < size/2 test here, unlike FilterInterpolation
__device__ __forceinline__ Corners corners(int w_i, int h_i, float fx, float fy, int W, int H)
{
    Corners c;
    const float x2 = __fadd_rn((float)w_i, fx), y2 = __fadd_rn((float)h_i, fy);
    c.in_range = x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1);
    c.L = (int)x2; c.T = (int)y2;
    c.R = min(c.L + 1, W - 1); c.Bm = min(c.T + 1, H - 1);
    return c;
}

// Splat, restructured.  Every in-range pixel adds the SAME triple (-d*fx, -d*fy, d) to the 2 x 2 block of
// cells {T, Bm} x {L, R} (:75-88).  Here the triple goes, with one vector RED, to cell (T, L) of the scratch image
// S; the box pass below then forms, with the reference's border behaviour (R = min(L+1, W-1), Bm = min(T+1, H-1):
// a block on the last column / row hits that column / row twice),
//   A[y][x] = sum_{dy,dx in {0,1}} wy(y,dy) * wx(x,dx) * S[y-dy][x-dx],   w(.,1) = 1,  wx(x,0) = (x == W-1 ? 2 : 1), wy alike.
// Where the splat takes its flow from.  lowres == 0: the full-resolution flow tensor [B,2,H,W].  lowres != 0 (SURVEY.md 8f
// rank 3): a quarter-resolution flow [B,2,H/4,W/4] that networks/DAIN.py:306-308 would first scale
// (`div_flow * temp * time_offset`, two fp32 multiplications) and then enlarge with nn.Upsample(scale_factor=4,
// mode='bilinear') (align_corners = False) into a full-resolution tensor: the same values are computed here on the fly,
// per pixel, from the four low-resolution neighbours (which stay in L1/L2) -- the full-resolution flow is never
// written or read.  Arithmetic: PyTorch's published upsample_bilinear2d (aten/src/ATen/native/UpSample.h:
// area_pixel_compute_source_index, cuda/UpSampleBilinear2d.cu), fp32, src = 0.25 * (dst + 0.5) - 0.5 clamped at 0.
struct FlowSource {
    const float *flow;
    int lowres, h, w;      // low-resolution extent when lowres != 0
    float s0, s1;          // the two scale factors, applied in this order to the low-resolution samples
};

__device__ __forceinline__ void upsample_coord(int dst, int in_size, int &i0, int &i1, float &l0, float &l1)
{
    float src = __fsub_rn(__fmul_rn(0.25f, __fadd_rn((float)dst, 0.5f)), 0.5f);
    src = src < 0.0f ? 0.0f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = __fsub_rn(src, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

__device__ __forceinline__ void load_flow(const FlowSource &fs, int b, int h_i, int w_i, int H, int W, float &fx, float &fy)
{
    if (!fs.lowres) {
        const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
        fx = ld_stream(fs.flow + ((size_t)b * 2 + 0) * HW + pix);
        fy = ld_stream(fs.flow + ((size_t)b * 2 + 1) * HW + pix);
        return;
    }
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    upsample_coord(h_i, fs.h, y0, y1, ly0, ly1);
    upsample_coord(w_i, fs.w, x0, x1, lx0, lx1);
    const size_t hw = (size_t)fs.h * fs.w;
    float v[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float *pl = fs.flow + ((size_t)b * 2 + ch) * hw;
        const float a = __fmul_rn(__fmul_rn(fs.s0, __ldg(pl + y0 * fs.w + x0)), fs.s1), bq = __fmul_rn(__fmul_rn(fs.s0, __ldg(pl + y0 * fs.w + x1)), fs.s1);
        const float c = __fmul_rn(__fmul_rn(fs.s0, __ldg(pl + y1 * fs.w + x0)), fs.s1), d = __fmul_rn(__fmul_rn(fs.s0, __ldg(pl + y1 * fs.w + x1)), fs.s1);
        v[ch] = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, a), __fmul_rn(lx1, bq))),
                          __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, c), __fmul_rn(lx1, d))));
    }
    fx = v[0]; fy = v[1];
}

// the enlarged flow as a tensor (tests; callers that need it next to the projection)
__global__ void __launch_bounds__(BX *BY)
flow_upsample4_kernel(FlowSource fs, float *__restrict__ out, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    float fx, fy;
    load_flow(fs, blockIdx.z, h_i, w_i, H, W, fx, fy);
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    out[((size_t)blockIdx.z * 2 + 0) * HW + pix] = fx;
    out[((size_t)blockIdx.z * 2 + 1) * HW + pix] = fy;
}

// Adjoint of the x4 enlargement (training through vfidkr_flowprojection_forward_lowres): the gradient with respect to a
// low-resolution sample is s0 * s1 * the sum over the full-resolution pixels whose bilinear footprint contains it, each
// with the weight upsample_coord gives it.  A pixel y reads rows i0 = floor(0.25 (y + 0.5) - 0.5) (clamped at 0) and
// i1 = i0 + 1 (clamped at h - 1), so row i is read by y in [4 i - 2, 4 i + 5] only: at most 8 x 8 terms per sample,
// gathered (no atomics, deterministic).
__global__ void __launch_bounds__(BX *BY)
flow_upsample4_backward_kernel(const float *__restrict__ gfull, float scale, float *__restrict__ glow, int h, int w)
{
    const int j = blockIdx.x * BX + threadIdx.x, i = blockIdx.y * BY + threadIdx.y;
    if (j >= w || i >= h) return;
    const int H = 4 * h, W = 4 * w;
    const int bc = blockIdx.z;     // batch item * 2 + flow component
    const float *g = gfull + (size_t)bc * H * W;
    float acc = 0.0f;
    for (int y = max(4 * i - 3, 0); y <= min(4 * i + 6, H - 1); ++y) {
        int y0, y1;
        float ly0, ly1;
        upsample_coord(y, h, y0, y1, ly0, ly1);
        const float wy = (y0 == i ? ly0 : 0.0f) + (y1 == i ? ly1 : 0.0f);
        if (wy == 0.0f) continue;
        float row = 0.0f;
        for (int x = max(4 * j - 3, 0); x <= min(4 * j + 6, W - 1); ++x) {
            int x0, x1;
            float lx0, lx1;
            upsample_coord(x, w, x0, x1, lx0, lx1);
            const float wx = (x0 == j ? lx0 : 0.0f) + (x1 == j ? lx1 : 0.0f);
            if (wx != 0.0f) row = fmaf(wx, __ldg(g + (size_t)y * W + x), row);
        }
        acc = fmaf(wy, row, acc);
    }
    glow[(size_t)bc * h * w + (size_t)i * w + j] = acc * scale;
}

// `clear` (may be null): the scratch image of the NEXT chunk of frames, zeroed here cell for cell (plain write-back
// stores: the lines stay in L2, where that chunk's REDs will find them).
template <bool DEPTH>
__global__ void __launch_bounds__(BX *BY)
projection_splat_kernel(const FlowSource fs, int b0, const float *__restrict__ depth, float4 *__restrict__ S,
                        float4 *__restrict__ clear, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    if (clear) clear[(size_t)b * HW + pix] = make_float4(0.f, 0.f, 0.f, 0.f);
    float fx, fy;
    load_flow(fs, b0 + b, h_i, w_i, H, W, fx, fy);   // b0: first frame of this chunk in the flow tensor
    const Corners c = corners(w_i, h_i, fx, fy, W, H);
    if (!c.in_range) return;
    const float d = DEPTH ? ld_stream(depth + (size_t)b * HW + pix) : 1.0f;
    const float vx = DEPTH ? -d * fx : -fx, vy = DEPTH ? -d * fy : -fy;   // :75-88 / depth :77-92
    atomicAdd(S + (size_t)b * HW + (size_t)c.T * W + c.L, make_float4(vx, vy, d, 0.0f));   // result unused -> REDG.F32x4
}

// Box pass + averaging (:128-135).  One warp owns a 32-column block of a row segment and walks it downwards,
// carrying the horizontally summed previous row in registers: S is read once (plus one halo row per segment and
// one halo column per block), count and output are written once, planar.
// column bitmap: for every column x a row of 64-bit words, word k = rows 64k .. 64k+63 (bit r = row 64k + r); the box
// pass writes it a byte (8 rows) at a time.  A hole n rows tall costs n / 64 loads per vertical scan.
__host__ __device__ inline int colwords(int H) { return (H + 63) >> 6; }
__device__ __forceinline__ unsigned char *colmask_byte(unsigned long long *cm, int x, int seg, int H)
{
    return reinterpret_cast<unsigned char *>(cm + (size_t)x * colwords(H) + (seg >> 3)) + (seg & 7);
}
// FIN_UNROLL rows of loads are requested before the first is used.  Four in flight per lane looked right for a DRAM-bound
// pass and was what round 1 shipped; measured with both libraries in one GPU call (profiles/r02/projection_finish_ab_v1.log)
// it is the registers that matter: 1 row = 32 registers = 64 resident warps per SM, 4 rows = 53 registers = 36 warps --
// DepthFlowProjection forward 364 -> 338 us (1: 338, 2: 350, 4: 364, 8: 412 us; 2 / 4 / 8 warps per block: no difference).
#ifndef VFIDKR_FIN_WARPS
#define VFIDKR_FIN_WARPS 4
#endif
#ifndef VFIDKR_FIN_UNROLL
#define VFIDKR_FIN_UNROLL 1
#endif
constexpr int FIN_ROWS = 8, FIN_WARPS = VFIDKR_FIN_WARPS, FIN_UNROLL = VFIDKR_FIN_UNROLL;   // short segments: one frame per launch must still fill 148 SMs

// raw loads of one row: this lane's cell and (lane 0 only) the cell left of the block
// CG: the scratch image was written earlier in the SAME launch by other SMs (fused pipeline kernel below): read it
// through L2 only (ld.global.cg) -- L1 and the read-only path are not coherent and may still hold the lines of the
// frame that used this buffer before.
template <bool CG = false>
__device__ __forceinline__ void load_row(const float4 *__restrict__ Srow, int x, int W, int lane, bool valid,
                                         float4 &cur, float4 &edge)
{
    cur = make_float4(0.f, 0.f, 0.f, 0.f);
    edge = cur;
    if (valid) {
        if (x < W) cur = CG ? __ldcg(Srow + x) : __ldcs(Srow + x);
        if (lane == 0 && x > 0 && x < W) edge = CG ? __ldcg(Srow + x - 1) : __ldg(Srow + x - 1);
    }
}
// horizontal half of the box: wx(x,0) * S[x] + S[x-1]
__device__ __forceinline__ float4 hsum_row(const float4 &cur, const float4 &edge, int x, int W, int lane)
{
    float4 left;
    left.x = __shfl_up_sync(0xffffffffu, cur.x, 1);
    left.y = __shfl_up_sync(0xffffffffu, cur.y, 1);
    left.z = __shfl_up_sync(0xffffffffu, cur.z, 1);
    if (lane == 0) left = edge;
    const float w0 = (x == W - 1) ? 2.0f : 1.0f;
    return make_float4(w0 * cur.x + left.x, w0 * cur.y + left.y, w0 * cur.z + left.z, 0.f);
}

__global__ void __launch_bounds__(32 * FIN_WARPS)
projection_finish_kernel(const float4 *__restrict__ S, float *__restrict__ count, float *__restrict__ out,
                         unsigned *__restrict__ rowmask, unsigned long long *__restrict__ colmask, unsigned *__restrict__ holemask,
                         int H, int W)
{
    // holemask (with rowmask / colmask): one bit per pixel, set where count <= 0 -- the holes; eight words (one per row)
    // per 32 x 8 block, contiguous, so that the hole filling finds its work without reading the count plane.
    // rowmask / colmask (inference only, else null): one bit per pixel, set where count != 0 -- the pixels the hole
    // filling may take values from (:175-213) -- packed along rows (32 columns per word) and along columns (the 8 rows
    // of this segment per byte), so that its scans step 32 columns / 8 rows per load instead of one pixel.
    static_assert(FIN_ROWS == 8, "colmask packs one box-pass segment per byte");
    const int lane = threadIdx.x, x = (blockIdx.x * FIN_WARPS + threadIdx.y) * 32 + lane;
    if ((blockIdx.x * FIN_WARPS + threadIdx.y) * 32 >= W) return;   // whole warp outside
    const int b = blockIdx.z, y0 = blockIdx.y * FIN_ROWS, y1 = min(y0 + FIN_ROWS, H);
    const size_t HW = (size_t)H * W;
    const float4 *Sb = S + (size_t)b * HW;
    float *ou = out + ((size_t)b * 2 + 0) * HW, *ov = ou + HW, *cn = count + (size_t)b * HW;
    float4 c0, e0;
    load_row(Sb + (size_t)max(y0 - 1, 0) * W, x, W, lane, y0 > 0, c0, e0);
    float4 prev = hsum_row(c0, e0, x, W, lane);
    unsigned colbits = 0, holeword = 0;
    for (int yb = y0; yb < y1; yb += FIN_UNROLL) {
        float4 cur[FIN_UNROLL], edge[FIN_UNROLL];
#pragma unroll
        for (int k = 0; k < FIN_UNROLL; ++k)   // all loads of the group are in flight before the first is used
            load_row(Sb + (size_t)min(yb + k, H - 1) * W, x, W, lane, yb + k < y1, cur[k], edge[k]);
#pragma unroll
        for (int k = 0; k < FIN_UNROLL; ++k) {
            const int y = yb + k;
            const float4 h = hsum_row(cur[k], edge[k], x, W, lane);
            const float w0 = (y == H - 1) ? 2.0f : 1.0f;
            float su = w0 * h.x + prev.x, sv = w0 * h.y + prev.y;
            const float sc = w0 * h.z + prev.z;
            prev = h;
            const bool live = y < y1 && x < W;
            if (live) {
                if (sc > 0.0f) { su = su / sc; sv = sv / sc; }   // :130-134
                const size_t a = (size_t)y * W + x;
                st_stream(cn + a, sc); st_stream(ou + a, su); st_stream(ov + a, sv);
            }
            if (rowmask) {   // uniform
                const bool src = live && sc != 0.0f;
                const unsigned m = __ballot_sync(0xffffffffu, src);
                if (lane == 0 && y < y1) rowmask[((size_t)b * H + y) * ((W + 31) >> 5) + (x >> 5)] = m;
                colbits |= (src ? 1u : 0u) << (y - y0);
                const unsigned hm = __ballot_sync(0xffffffffu, live && !(sc > 0.0f));   // count <= 0 (:171)
                if (lane == (y - y0)) holeword = hm;
            }
        }
    }
    if (colmask && x < W) *colmask_byte(colmask + (size_t)b * W * colwords(H), x, blockIdx.y, H) = (unsigned char)colbits;
    if (holemask && lane < FIN_ROWS)
        holemask[(((size_t)b * ((H + 7) >> 3) + blockIdx.y) * ((W + 31) >> 5) + (x >> 5)) * FIN_ROWS + lane] = holeword;
}

// hole filling (:171-232).  Reads only non-hole pixels (count != 0), which this kernel never writes, so running it in
// place is race-free.  The reference walks from every hole pixel one pixel at a time in four directions, unbounded;
// here the walks run over the source bitmaps written by the box pass: a word covers 32 columns, a byte 8 rows, so a hole
// n pixels wide costs n / 32 (n / 8) loads per direction instead of n, and the nearest source inside a word is a
// clz / ffs.  Each scan returns the distance to the nearest source (0: none before the edge of the plane).
template <bool CG, typename T>
__device__ __forceinline__ unsigned ldm(const T *p) { return CG ? (unsigned)__ldcg(p) : (unsigned)__ldg(p); }
template <bool CG = false>
__device__ __forceinline__ int scan_left(const unsigned *__restrict__ rm, int x)
{
    int wi = x >> 5;
    unsigned m = ldm<CG>(rm + wi) & ((1u << (x & 31)) - 1u);
    for (;;) {
        if (m) return x - ((wi << 5) + 31 - __clz(m));
        if (--wi < 0) return 0;
        m = ldm<CG>(rm + wi);
    }
}
template <bool CG = false>
__device__ __forceinline__ int scan_right(const unsigned *__restrict__ rm, int x, int WW)
{
    int wi = x >> 5;
    unsigned m = ldm<CG>(rm + wi) & ~((2u << (x & 31)) - 1u);   // bits above x (none when x & 31 == 31)
    for (;;) {
        if (m) return (wi << 5) + __ffs(m) - 1 - x;
        if (++wi >= WW) return 0;
        m = ldm<CG>(rm + wi);
    }
}
template <bool CG = false>
__device__ __forceinline__ int scan_up(const unsigned long long *__restrict__ col, int y)
{
    int wi = y >> 6;
    unsigned long long m = (CG ? __ldcg(col + wi) : __ldg(col + wi)) & ((1ull << (y & 63)) - 1ull);
    for (;;) {
        if (m) return y - ((wi << 6) + 63 - __clzll((long long)m));
        if (--wi < 0) return 0;
        m = CG ? __ldcg(col + wi) : __ldg(col + wi);
    }
}
template <bool CG = false>
__device__ __forceinline__ int scan_down(const unsigned long long *__restrict__ col, int y, int HWd)
{
    int wi = y >> 6;
    unsigned long long m = (CG ? __ldcg(col + wi) : __ldg(col + wi)) & ~((2ull << (y & 63)) - 1ull);   // bits above y (none when y & 63 == 63)
    for (;;) {
        if (m) return (wi << 6) + __ffsll((long long)m) - 1 - y;
        if (++wi >= HWd) return 0;
        m = CG ? __ldcg(col + wi) : __ldg(col + wi);
    }
}

// One hole (:175-232): the nearest source pixel in each of the four axis directions, found on the bitmaps; the hole
// becomes the mean of those that exist.  cnb / ob / rm / cm: count plane, the two output planes, row and column bitmaps of
// ONE frame.  CG: everything was written earlier in the same launch by other SMs -> L2-coherent loads.
template <bool CG>
__device__ __forceinline__ void fill_one(const float *__restrict__ cnb, float *__restrict__ ob, const unsigned *__restrict__ rm,
                                         const unsigned long long *__restrict__ cm, int x, int y, int H, int W)
{
    auto ld = [](const float *p) { return CG ? __ldcg(p) : __ldg(p); };
    const size_t HW = (size_t)H * W;
    const int WW = (W + 31) >> 5, HWd = colwords(H);
    const float *cn = cnb + (size_t)y * W + x;
    const int dl = scan_left<CG>(rm + (size_t)y * WW, x);
    const int dr = scan_right<CG>(rm + (size_t)y * WW, x, WW);
    const int du = scan_up<CG>(cm + (size_t)x * HWd, y);
    const int dd = scan_down<CG>(cm + (size_t)x * HWd, y, HWd);
    // the counts the reference's loops end on (0 when a scan ran off the plane)
    const float lt = dl ? ld(cn - dl) : 0.0f, rt = dr ? ld(cn + dr) : 0.0f;
    const float ut = du ? ld(cn - (long long)du * W) : 0.0f, dt = dd ? ld(cn + (long long)dd * W) : 0.0f;
    if (lt + rt + ut + dt <= 0.0f) return;
    const float l = lt > 0.0f ? 1.f : 0.f, r = rt > 0.0f ? 1.f : 0.f;
    const float u = ut > 0.0f ? 1.f : 0.f, d = dt > 0.0f ? 1.f : 0.f;
    const float den = l + r + u + d;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float *o = ob + (size_t)ch * HW + (size_t)y * W + x;
        // the sources are non-hole pixels, final since the averaging pass and never written here (plain loads in the
        // three-kernel path: no read-only path for a buffer this kernel also writes)
        const float v = CG ? l * __ldcg(o - dl) + r * __ldcg(o + dr) + u * __ldcg(o - (long long)du * W) + d * __ldcg(o + (long long)dd * W)
                           : l * o[-dl] + r * o[dr] + u * o[-(long long)du * W] + d * o[(long long)dd * W];
        *o = v / den;
    }
}

// Hole filling, driven by the hole bitmask the box pass wrote (eight words per 32 x 8 block): warps stride over the
// blocks of the whole batch, skip blocks and rows without holes after one 32-byte load, and a lane fills the hole in its
// column.  Neighbouring holes are filled by neighbouring lanes / consecutive rows, so their scans and gathers share
// sectors.  (Round 1 launched one CTA per 32 x 8 tile that re-read the count plane and compacted its holes: 71 k CTAs
// at 1080p x 8, 75-180 us of CTA turnover for a few per cent of hole pixels.)
constexpr int FILL_WARPS = 8;
// eight resident blocks (32 registers, no spills) instead of the five that 46 registers allow: the scans are dependent
// loads, more warps hide them (up4 flow: 388.6 -> 375.7 us for the whole DepthFlowProjection forward; scene 338.6 -> 336.6)
#ifndef VFIDKR_FILL_MINB
#define VFIDKR_FILL_MINB 8
#endif
__global__ void __launch_bounds__(32 * FILL_WARPS, VFIDKR_FILL_MINB)
projection_fill_mask_kernel(const float *__restrict__ count, float *__restrict__ out, const unsigned *__restrict__ rowmask,
                            const unsigned long long *__restrict__ colmask, const unsigned *__restrict__ holemask,
                            int B, int H, int W, const FastDiv div_items, const FastDiv div_bw)
{
    const int lane = threadIdx.x & 31;
    const int bw = (W + 31) >> 5, HB = (H + 7) >> 3, nitems = bw * HB;
    const size_t HW = (size_t)H * W;
    const int total = B * nitems;     // < 2^31 (launcher)
    for (int g = blockIdx.x * FILL_WARPS + (threadIdx.x >> 5); g < total; g += gridDim.x * FILL_WARPS) {
        const unsigned mine = lane < FIN_ROWS ? __ldg(holemask + (size_t)g * FIN_ROWS + lane) : 0u;
        if (!__any_sync(0xffffffffu, mine != 0u)) continue;
        const int f = div_items.quot(g), it = g - f * nitems;
        const int seg = div_bw.quot(it), x0 = (it - seg * bw) * 32;
        // the block's holes in raster order, dealt to the lanes densely: hole n is the (n - before[r])-th set bit of row r
        unsigned w[FIN_ROWS];
        int upto[FIN_ROWS], n_holes = 0;
#pragma unroll
        for (int k = 0; k < FIN_ROWS; ++k) {
            w[k] = __shfl_sync(0xffffffffu, mine, k);
            n_holes += __popc(w[k]);
            upto[k] = n_holes;          // holes in rows 0 .. k
        }
        for (int n = lane; n < n_holes; n += 32) {
            int r = 0, before = 0;
            unsigned word = w[0];
#pragma unroll
            for (int k = 1; k < FIN_ROWS; ++k)
                if (n >= upto[k - 1]) { r = k; before = upto[k - 1]; word = w[k]; }
            const int bit = (int)__fns(word, 0, n - before + 1);
            fill_one<false>(count + (size_t)f * HW, out + (size_t)f * 2 * HW, rowmask + (size_t)f * H * bw,
                            colmask + (size_t)f * W * colwords(H), x0 + bit, seg * FIN_ROWS + r, H, W);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// What else was built for the forward in round 2, verified against the oracle, measured and REMOVED (git history;
// profiles/r02/time_projection_v*.log, ncu_dproj_chunks_v1.txt, launches_dproj_chunks_warm_v1.csv) -- the whole-batch
// splat + box pass above stayed the fastest at 1080p x 8 (313 us without hole filling):
//   * one cooperative kernel, splat workers and box-pass workers on different frames, per-frame counters in global
//     memory: every hand-off (fence, counter, poll) costs ~2 us and a frame needs six in sequence -- 339-736 us;
//   * chunks of one frame with two alternating scratch images meant to stay in L2, persistent kernels (one and four
//     pixels per thread), programmatic dependent launch: 316-350 us.  The premise failed: with a 36.5 MB image being
//     RED into, a second one being cleared and 54 MB of streams per frame, L2 does NOT keep the image (ncu, warm caches:
//     the box pass reads its 36.6 MB from DRAM again, L2 hit rate 47-50 %), and a one-frame launch takes 19-28 us however
//     few instructions it issues (3.3 M vs 7.4 M warp-instructions per frame made no difference): latency-bound at
//     32 warps per SM.  The microbenchmark's 360 G vector REDs/s into a resident 36 MB image (tools/microbench/ffma2.cu)
//     needs the image to be the ONLY thing in flight.
// What did pay: the hole filling below (bitmask-driven, dense lanes, 64-row column words).
// ---------------------------------------------------------------------------------------------------------------

// backward gather (:266-297; depth :276-337)
template <bool DEPTH>
__global__ void __launch_bounds__(BX *BY)
projection_backward_kernel(const float *__restrict__ flow, const float *__restrict__ depth,
                           const float *__restrict__ count, const float *__restrict__ out,
                           const float *__restrict__ gout, float *__restrict__ gi1, float *__restrict__ gi2,
                           int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flow + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flow + ((size_t)b * 2 + 1) * HW + pix);
    const Corners c = corners(w_i, h_i, fx, fy, W, H);
    float su = 0.0f, sv = 0.0f, sd = 0.0f;
    if (c.in_range) {
        const float d = DEPTH ? ld_stream(depth + (size_t)b * HW + pix) : 1.0f;
        const float *gu = gout + ((size_t)b * 2 + 0) * HW, *gv = gu + HW, *cn = count + (size_t)b * HW;
        const float *ou = DEPTH ? out + ((size_t)b * 2 + 0) * HW : nullptr;
        const int a[4] = {c.T * W + c.L, c.T * W + c.R, c.Bm * W + c.L, c.Bm * W + c.R};   // 32-bit: H * W < 2^31 (launcher)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // one IEEE division per corner (1 / count) instead of the reference's four: products differ from its
            // quotients by at most an ulp or two, far inside the gradient tolerance
            const float rc = 1.0f / __ldg(cn + a[k]);
            const float gU = __ldg(gu + a[k]) * rc, gV = __ldg(gv + a[k]) * rc;   // gradout / count
            if (DEPTH) {
                su += -gU * d;                                         // depth :289-296
                sv += -gV * d;
                sd += -gU * (fx - __ldg(ou + a[k]));                   // depth :311-322
                sd += -gV * (fy - __ldg(ou + HW + a[k]));              // depth :324-335
            } else {
                su += -gU;                                             // :277-284
                sv += -gV;
            }
        }
    }
    st_stream(gi1 + ((size_t)b * 2 + 0) * HW + pix, su);
    st_stream(gi1 + ((size_t)b * 2 + 1) * HW + pix, sv);
    if (DEPTH) st_stream(gi2 + (size_t)b * HW + pix, sd);
}


// ---------------------------------------------------------------------------------------------------------------
// MinDepthFlowProjection (my_package/MinDepthFlowProjection/mindepthflowprojection_cuda_kernel.cu:29-312), SURVEY.md 8f
// rank 4.  Intended semantics of the reference: every in-range source pixel competes for ONE cell, the top-left one
// (T, L) (the other three corners are commented out, :86-113); the source with the largest input2 (inverse depth:
// the closest surface) wins, provided it is > 0 (count starts at 0, :79); output = -flow of the winner, count = its
// input2.  The reference does this with a non-atomic read-compare-write (:79-84), so its result depends on thread
// timing even without ties.  Here the competition is a 64-bit atomicMax on (bits of input2) << 32 | ~pixel index:
// the largest input2 wins and, among equals, the lowest pixel index -- deterministic, and equal to what the reference
// computes whenever its race does not strike.  A second pass decodes the winners and writes the hole-filling bitmaps.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX *BY)
mindepth_select_kernel(const float *__restrict__ flow, const float *__restrict__ depth, unsigned long long *__restrict__ keys,
                       int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flow + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flow + ((size_t)b * 2 + 1) * HW + pix);
    const Corners c = corners(w_i, h_i, fx, fy, W, H);
    if (!c.in_range) return;
    const float d = ld_stream(depth + (size_t)b * HW + pix);
    if (!(d > 0.0f)) return;   // temp > old_exist with old_exist >= 0 (:79)
    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (0xffffffffu - (unsigned)pix);
    atomicMax(keys + (size_t)b * HW + (size_t)c.T * W + c.L, key);   // positive floats order like their bit patterns
}

__global__ void __launch_bounds__(32 * FIN_WARPS)
mindepth_resolve_kernel(const unsigned long long *__restrict__ keys, const float *__restrict__ flow, float *__restrict__ count,
                        float *__restrict__ out, unsigned *__restrict__ rowmask, unsigned long long *__restrict__ colmask,
                        unsigned *__restrict__ holemask, int H, int W)
{
    const int lane = threadIdx.x, x = (blockIdx.x * FIN_WARPS + threadIdx.y) * 32 + lane;
    if ((blockIdx.x * FIN_WARPS + threadIdx.y) * 32 >= W) return;   // whole warp outside
    const int b = blockIdx.z, y0 = blockIdx.y * FIN_ROWS, y1 = min(y0 + FIN_ROWS, H);
    const size_t HW = (size_t)H * W;
    const float *fu = flow + (size_t)b * 2 * HW, *fv = fu + HW;
    unsigned colbits = 0, holeword = 0;
    for (int y = y0; y < y0 + FIN_ROWS; ++y) {
        const bool live = y < y1 && x < W;
        float cnt = 0.0f, u = 0.0f, v = 0.0f;
        if (live) {
            const size_t a = (size_t)y * W + x;
            const unsigned long long key = __ldcs(keys + (size_t)b * HW + a);
            if (key != 0ull) {
                const unsigned src = 0xffffffffu - (unsigned)(key & 0xffffffffull);
                cnt = __uint_as_float((unsigned)(key >> 32));
                u = -__ldg(fu + src);   // :80-81
                v = -__ldg(fv + src);
            }
            st_stream(count + (size_t)b * HW + a, cnt);
            st_stream(out + (size_t)b * 2 * HW + a, u);
            st_stream(out + (size_t)b * 2 * HW + HW + a, v);
        }
        if (rowmask) {   // uniform
            const bool src_ok = live && cnt != 0.0f;
            const unsigned m = __ballot_sync(0xffffffffu, src_ok);
            if (lane == 0 && y < y1) rowmask[((size_t)b * H + y) * ((W + 31) >> 5) + (x >> 5)] = m;
            colbits |= (src_ok ? 1u : 0u) << (y - y0);
            const unsigned hm = __ballot_sync(0xffffffffu, live && !(cnt > 0.0f));
            if (lane == y - y0) holeword = hm;
        }
    }
    if (colmask && x < W) *colmask_byte(colmask + (size_t)b * W * colwords(H), x, blockIdx.y, H) = (unsigned char)colbits;
    if (holemask && lane < FIN_ROWS)
        holemask[(((size_t)b * ((H + 7) >> 3) + blockIdx.y) * ((W + 31) >> 5) + (x >> 5)) * FIN_ROWS + lane] = holeword;
}

// backward (:216-312): a source pixel receives -gradoutput of each of its four corners whose count equals its input2
// (all four are tested although the forward only writes the top-left one -- as written)
__global__ void __launch_bounds__(BX *BY)
mindepth_backward_kernel(const float *__restrict__ flow, const float *__restrict__ depth, const float *__restrict__ count,
                         const float *__restrict__ gout, float *__restrict__ gi1, float *__restrict__ gi2, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flow + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flow + ((size_t)b * 2 + 1) * HW + pix);
    const Corners c = corners(w_i, h_i, fx, fy, W, H);
    float su = 0.0f, sv = 0.0f;
    if (c.in_range) {
        const float d = ld_stream(depth + (size_t)b * HW + pix);
        const float *gu = gout + (size_t)b * 2 * HW, *gv = gu + HW, *cn = count + (size_t)b * HW;
        const int a[4] = {c.T * W + c.L, c.T * W + c.R, c.Bm * W + c.L, c.Bm * W + c.R};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (d == __ldg(cn + a[k])) { su += -__ldg(gu + a[k]); sv += -__ldg(gv + a[k]); }
    }
    st_stream(gi1 + ((size_t)b * 2 + 0) * HW + pix, su);
    st_stream(gi1 + ((size_t)b * 2 + 1) * HW + pix, sv);
    st_stream(gi2 + (size_t)b * HW + pix, 0.0f);   // the reference never writes gradinput2: it stays the caller's zeros
}

}  // namespace

namespace {

// the hole-filling bitmaps of a batch inside one scratch block: [column bitmap (64-bit words) | row bitmap | hole bitmask]
struct FillMaps {
    unsigned long long *colmask; unsigned *rowmask, *holemask;
    size_t colmask_bytes, bytes;
};
static FillMaps fill_maps(char *base, int B, int H, int W, bool enabled)
{
    FillMaps m{nullptr, nullptr, nullptr, 0, 0};
    if (!enabled) return m;
    const size_t bw = ((size_t)W + 31) >> 5, HB = ((size_t)H + 7) >> 3;
    m.colmask_bytes = sizeof(unsigned long long) * B * W * colwords(H);
    const size_t rowmask_bytes = sizeof(unsigned) * B * H * bw, holemask_bytes = sizeof(unsigned) * B * HB * bw * FIN_ROWS;
    m.bytes = m.colmask_bytes + rowmask_bytes + holemask_bytes;
    if (base) {
        m.colmask = reinterpret_cast<unsigned long long *>(base);
        m.rowmask = reinterpret_cast<unsigned *>(base + m.colmask_bytes);
        m.holemask = reinterpret_cast<unsigned *>(base + m.colmask_bytes + rowmask_bytes);
    }
    return m;
}

// launch of the mask-walking hole filling for a whole batch (shared by every forward path)
static int launch_fill(const float *count, float *out, const FillMaps &m, int B, int H, int W, cudaStream_t s)
{
    const int bw = (W + 31) >> 5, HB = (H + 7) >> 3;
    const long long total = (long long)B * bw * HB;
    if (total >= (1ll << 31)) return VFIDKR_ERR_ARG;
    const unsigned nb = (unsigned)std::min<long long>((total + FILL_WARPS - 1) / FILL_WARPS, (long long)sm_count() * 8);
    projection_fill_mask_kernel<<<nb, 32 * FILL_WARPS, 0, s>>>(count, out, m.rowmask, m.colmask, m.holemask, B, H, W,
                                                              FastDiv((unsigned)(bw * HB)), FastDiv((unsigned)bw));
    note_launch();
    return check_launch("flow projection hole filling");
}

template <bool DEPTH>
int projection_forward(const FlowSource fs, const float *depth, float *count, float *out,
                       int B, int H, int W, int fillhole, cudaStream_t s)
{
    const float *flow = fs.flow;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !flow || !count || !out || (DEPTH && !depth)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31) || ceil_div(H, FIN_ROWS) > 65535u) return VFIDKR_ERR_ARG;
    const size_t HW = (size_t)H * W;
    // ONE splat and ONE box pass over the whole batch (one pixel / one 32 x 8 block per thread / warp), then the hole filling
    const size_t scratch_bytes = sizeof(float4) * B * HW;
    void *scratch = nullptr;
    int e = stream_scratch_alloc(&scratch, scratch_bytes + fill_maps(nullptr, B, H, W, fillhole).bytes, s);
    if (e) return e;
    float4 *S = static_cast<float4 *>(scratch);
    const FillMaps fm = fill_maps(static_cast<char *>(scratch) + scratch_bytes, B, H, W, fillhole);
    e = set_error(cudaMemsetAsync(S, 0, scratch_bytes, s), "clear projection scratch");
    if (!e && fillhole) e = set_error(cudaMemsetAsync(fm.colmask, 0, fm.colmask_bytes, s), "clear column bitmap");
    if (!e) {
        dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
        projection_splat_kernel<DEPTH><<<grid, block, 0, s>>>(fs, 0, depth, S, nullptr, H, W);
        dim3 fblock(32, FIN_WARPS), fgrid(ceil_div(W, 32 * FIN_WARPS), ceil_div(H, FIN_ROWS), B);
        projection_finish_kernel<<<fgrid, fblock, 0, s>>>(S, count, out, fm.rowmask, fm.colmask, fm.holemask, H, W);
        note_launch(2);
        e = check_launch("flow projection forward");
        if (!e && fillhole) e = launch_fill(count, out, fm, B, H, W, s);
    }
    const int e2 = set_error(cudaFreeAsync(scratch, s), "projection scratch (cudaFreeAsync)");
    return e ? e : e2;
}

template <bool DEPTH>
int projection_backward(const float *flow, const float *depth, const float *count, const float *out,
                        const float *gout, float *gi1, float *gi2, int B, int H, int W, cudaStream_t s)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !flow || !count || !gout || !gi1) return VFIDKR_ERR_ARG;
    if (DEPTH && (!depth || !out || !gi2)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    projection_backward_kernel<DEPTH><<<grid, block, 0, s>>>(flow, depth, count, out, gout, gi1, gi2, H, W);
    note_launch();
    return check_launch("flow projection backward");
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_flowprojection_forward(const float *input1, float *count, float *output,
                                             int B, int H, int W, int fillhole, vfidkr_stream_t s)
{ return projection_forward<false>(FlowSource{input1, 0, 0, 0, 1.f, 1.f}, nullptr, count, output, B, H, W, fillhole, (cudaStream_t)s); }

VFIDKR_API int vfidkr_flowprojection_backward(const float *input1, const float *count, const float *gradoutput,
                                              float *gradinput1, int B, int H, int W, vfidkr_stream_t s)
{ return projection_backward<false>(input1, nullptr, count, nullptr, gradoutput, gradinput1, nullptr, B, H, W, (cudaStream_t)s); }

VFIDKR_API int vfidkr_depthflowprojection_forward(const float *input1, const float *input2, float *count,
                                                  float *output, int B, int H, int W, int fillhole,
                                                  vfidkr_stream_t s)
{ return projection_forward<true>(FlowSource{input1, 0, 0, 0, 1.f, 1.f}, input2, count, output, B, H, W, fillhole, (cudaStream_t)s); }

VFIDKR_API int vfidkr_depthflowprojection_backward(const float *input1, const float *input2, const float *count,
                                                   const float *output, const float *gradoutput,
                                                   float *gradinput1, float *gradinput2,
                                                   int B, int H, int W, vfidkr_stream_t s)
{ return projection_backward<true>(input1, input2, count, output, gradoutput, gradinput1, gradinput2, B, H, W, (cudaStream_t)s); }

VFIDKR_API int vfidkr_mindepthflowprojection_forward(const float *input1, const float *input2, float *count, float *output,
                                                     int B, int H, int W, int fillhole, vfidkr_stream_t stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !input1 || !input2 || !count || !output) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31) || ceil_div(H, FIN_ROWS) > 65535u) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t HW = (size_t)H * W;
    const size_t key_bytes = sizeof(unsigned long long) * B * HW;
    void *mem = nullptr;
    int e = stream_scratch_alloc(&mem, key_bytes + fill_maps(nullptr, B, H, W, fillhole).bytes, s);
    if (e) return e;
    unsigned long long *keys = static_cast<unsigned long long *>(mem);
    const FillMaps fm = fill_maps(static_cast<char *>(mem) + key_bytes, B, H, W, fillhole);
    if (fillhole) e = set_error(cudaMemsetAsync(fm.colmask, 0, fm.colmask_bytes, s), "clear column bitmap");
    if (e) { cudaFreeAsync(mem, s); return e; }
    e = set_error(cudaMemsetAsync(keys, 0, key_bytes, s), "clear min-depth keys");
    if (!e) {
        dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
        mindepth_select_kernel<<<grid, block, 0, s>>>(input1, input2, keys, H, W);
        dim3 fblock(32, FIN_WARPS), fgrid(ceil_div(W, 32 * FIN_WARPS), ceil_div(H, FIN_ROWS), B);
        mindepth_resolve_kernel<<<fgrid, fblock, 0, s>>>(keys, input1, count, output, fm.rowmask, fm.colmask, fm.holemask, H, W);
        note_launch(2);
        e = check_launch("min-depth flow projection forward");
        if (!e && fillhole) e = launch_fill(count, output, fm, B, H, W, s);
    }
    const int e2 = set_error(cudaFreeAsync(mem, s), "min-depth scratch (cudaFreeAsync)");
    return e ? e : e2;
}

VFIDKR_API int vfidkr_mindepthflowprojection_backward(const float *input1, const float *input2, const float *count,
                                                      const float *gradoutput, float *gradinput1, float *gradinput2,
                                                      int B, int H, int W, vfidkr_stream_t stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !input1 || !input2 || !count || !gradoutput || !gradinput1 || !gradinput2)
        return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    mindepth_backward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(input1, input2, count, gradoutput, gradinput1, gradinput2, H, W);
    note_launch();
    return check_launch("min-depth flow projection backward");
}

// the projections fed by a quarter-resolution flow (see FlowSource).  h, w: low-resolution extent;
// the outputs are [B,*,4h,4w].  input2 (inverse depth, full resolution) may be null: FlowProjection.
VFIDKR_API int vfidkr_flowprojection_forward_lowres(const float *flow_lowres, float scale0, float scale1, const float *input2,
                                                    float *count, float *output, int B, int h, int w, int fillhole,
                                                    vfidkr_stream_t s)
{
    if (h <= 0 || w <= 0 || h > (1 << 20) || w > (1 << 20)) return VFIDKR_ERR_ARG;
    const FlowSource fs{flow_lowres, 1, h, w, scale0, scale1};
    return input2 ? projection_forward<true>(fs, input2, count, output, B, 4 * h, 4 * w, fillhole, (cudaStream_t)s)
                  : projection_forward<false>(fs, nullptr, count, output, B, 4 * h, 4 * w, fillhole, (cudaStream_t)s);
}

VFIDKR_API int vfidkr_flow_upsample4_backward(const float *grad_output, float scale0, float scale1, float *grad_flow_lowres,
                                              int B, int h, int w, vfidkr_stream_t s)
{
    if (B <= 0 || 2 * B > 65535 || h <= 0 || w <= 0 || h > (1 << 20) || w > (1 << 20) || !grad_output || !grad_flow_lowres) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(w, BX), ceil_div(h, BY), 2 * B);
    flow_upsample4_backward_kernel<<<grid, block, 0, (cudaStream_t)s>>>(grad_output, scale0 * scale1, grad_flow_lowres, h, w);
    note_launch();
    return check_launch("flow upsample x4 backward");
}

VFIDKR_API int vfidkr_flow_upsample4(const float *flow_lowres, float scale0, float scale1, float *output, int B, int h, int w,
                                     vfidkr_stream_t s)
{
    if (B <= 0 || B > 65535 || h <= 0 || w <= 0 || h > (1 << 20) || w > (1 << 20) || !flow_lowres || !output) return VFIDKR_ERR_ARG;
    const int H = 4 * h, W = 4 * w;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    flow_upsample4_kernel<<<grid, block, 0, (cudaStream_t)s>>>(FlowSource{flow_lowres, 1, h, w, scale0, scale1}, output, H, W);
    note_launch();
    return check_launch("flow upsample x4");
}
