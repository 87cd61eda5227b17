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
constexpr int FIN_ROWS = 8, FIN_WARPS = 4, FIN_UNROLL = 4;   // short segments: one frame per launch must still fill 148 SMs

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
                         unsigned *__restrict__ rowmask, unsigned char *__restrict__ colmask, int H, int W)
{
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
    unsigned colbits = 0;
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
            }
        }
    }
    if (colmask && x < W) colmask[((size_t)b * ((H + 7) >> 3) + blockIdx.y) * W + x] = (unsigned char)colbits;
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
__device__ __forceinline__ int scan_up(const unsigned char *__restrict__ cm, int y, int W)
{
    int bi = y >> 3;
    unsigned m = ldm<CG>(cm + (size_t)bi * W) & ((1u << (y & 7)) - 1u);
    for (;;) {
        if (m) return y - ((bi << 3) + 31 - __clz(m));
        if (--bi < 0) return 0;
        m = ldm<CG>(cm + (size_t)bi * W);
    }
}
template <bool CG = false>
__device__ __forceinline__ int scan_down(const unsigned char *__restrict__ cm, int y, int W, int HB)
{
    int bi = y >> 3;
    unsigned m = ldm<CG>(cm + (size_t)bi * W) & 0xffu & ~((2u << (y & 7)) - 1u);
    for (;;) {
        if (m) return (bi << 3) + __ffs(m) - 1 - y;
        if (++bi >= HB) return 0;
        m = ldm<CG>(cm + (size_t)bi * W);
    }
}

// One hole (:175-232): the nearest source pixel in each of the four axis directions, found on the bitmaps; the hole
// becomes the mean of those that exist.  cnb / ob / rm / cm: count plane, the two output planes, row and column bitmaps of
// ONE frame.  CG: everything was written earlier in the same launch by other SMs -> L2-coherent loads.
template <bool CG>
__device__ __forceinline__ void fill_one(const float *__restrict__ cnb, float *__restrict__ ob, const unsigned *__restrict__ rm,
                                         const unsigned char *__restrict__ cm, int x, int y, int H, int W)
{
    auto ld = [](const float *p) { return CG ? __ldcg(p) : __ldg(p); };
    const size_t HW = (size_t)H * W;
    const int WW = (W + 31) >> 5, HB = (H + 7) >> 3;
    const float *cn = cnb + (size_t)y * W + x;
    const int dl = scan_left<CG>(rm + (size_t)y * WW, x);
    const int dr = scan_right<CG>(rm + (size_t)y * WW, x, WW);
    const int du = scan_up<CG>(cm + x, y, W);
    const int dd = scan_down<CG>(cm + x, y, W, HB);
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

// Holes are sparse (a few per cent of the pixels) but scattered, so with one thread per pixel most warps would run
// the scans for one or two live lanes.  The CTA therefore first COMPACTS its holes into a shared list (ballot +
// one shared atomic per warp) and then fills them with dense warps.
__global__ void __launch_bounds__(BX *BY)
projection_fillhole_kernel(const float *__restrict__ count, float *__restrict__ out, const unsigned *__restrict__ rowmask,
                           const unsigned char *__restrict__ colmask, int H, int W)
{
    __shared__ int s_n;
    __shared__ unsigned short s_list[BX * BY];
    const int tid = threadIdx.y * BX + threadIdx.x, lane = threadIdx.x;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int WW = (W + 31) >> 5, HB = (H + 7) >> 3;
    const float *cnb = count + (size_t)b * HW;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const bool hole = w_i < W && h_i < H && !(__ldcs(cnb + (size_t)h_i * W + w_i) > 0.0f);   // count <= 0 (:171)
    const unsigned m = __ballot_sync(0xffffffffu, hole);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hole) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)tid;
    }
    __syncthreads();
    const int n = s_n;
    for (int q = tid; q < n; q += BX * BY) {
        const int t = s_list[q];
        fill_one<false>(cnb, out + (size_t)b * 2 * HW, rowmask + (size_t)b * H * WW, colmask + (size_t)b * HB * W,
                        blockIdx.x * BX + (t & (BX - 1)), blockIdx.y * BY + t / BX, H, W);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused forward for batches: ONE persistent, cooperatively launched kernel pipelines the frames of the batch.
//
// The splat is bound by the rate at which L2 retires vector REDs, the box pass by L2 / DRAM bandwidth: different units,
// so the two run CONCURRENTLY on different frames.  Half of the CTAs (one per SM) are splat workers, the other half
// (one more per SM) are finish workers; each role walks the frames in order with a static share of the frame's tiles,
// and the roles meet only through per-frame counters in global memory:
//     splat(f)   needs the scratch image f % NB clean      (c_done[f - NB] == workers; the first NB images: see below)
//     finish(f)  needs every splat worker done with f      (s_done[f] == workers)
//     clear(f)   needs every finish worker done with f     (f_done[f] == workers) -- it is run one frame LATE, after
//                finish(f + 1), when that condition has long been true: no worker ever waits at a barrier
// so the splat workers run up to NB - 1 frames ahead and the RED unit never waits for a box pass.  NB = 3 scratch
// images of 16 B per pixel rotate; at 1080p they and the streams around them live in the 126 MB L2, so the scratch
// makes no DRAM round trip (the three-kernel path below moves 2.9x the algorithmic bytes).  Image 0 is cleared on the
// stream before the launch, images 1 .. NB-1 by the finish workers while the first splat runs.
// Holes are appended to ONE list for the batch by the box pass (one global atomic per 32 x 4 block that has any); a
// second, small launch fills exactly those pixels.  (A first version filled the holes of frame f inside the pipeline,
// behind a barrier of the finish workers: its chain of dependent L2 round trips -- barrier, bitmap scans, value loads --
// cost 30-90 us PER FRAME with only one frame's holes in flight; measured 736 us against 421 us for three kernels.)
// All cross-SM traffic inside the launch uses L2-coherent accesses (RED, ld.global.cg / st.global.cg): L1 is not
// coherent and the images are reused.  The counters need every CTA resident: cudaLaunchCooperativeKernel guarantees it
// (or fails -> three-kernel path).
// ---------------------------------------------------------------------------------------------------------------
namespace pipe {

constexpr int NT = 512;            // threads per CTA, both roles
constexpr int NB_MAX = 3;          // scratch images in rotation
constexpr int SPLAT_UNROLL = 4;    // 32 x 16 pixel sub-tiles whose loads a splat worker keeps in flight
constexpr int WARPS = NT / 32;

struct FrameCtl { unsigned s_done, f_done, c_done, pad; };        // zeroed on the stream before the launch
struct Ctl { unsigned pre_done, h_count, pad[2]; };                // images 1 .. NB-1 cleared; holes listed so far

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// every thread of the CTA calls these two
__device__ __forceinline__ void cta_wait(const unsigned *flag, unsigned target)
{
    if (threadIdx.x == 0) {
        unsigned ns = 32;
        while (ld_acquire(flag) < target) { __nanosleep(ns); if (ns < 512) ns <<= 1; }
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void cta_signal(unsigned *flag)
{
    __syncthreads();   // every thread's REDs / stores are issued ...
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(flag, 1u); }   // ... and ordered before the count (cumulative fence)
}

template <bool DEPTH>
__device__ __forceinline__ void splat_frame(const FlowSource &fs, const float *__restrict__ depth_f, int frame,
                                            float4 *__restrict__ S, int H, int W, int worker, int nworkers, const FastDiv div_tx)
{
    const int tiles_x = (W + 31) >> 5, ntiles = tiles_x * ((H + 15) >> 4);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    for (int t0 = worker; t0 < ntiles; t0 += nworkers * SPLAT_UNROLL) {
        float fx[SPLAT_UNROLL], fy[SPLAT_UNROLL], d[SPLAT_UNROLL];
        int wi[SPLAT_UNROLL], hi[SPLAT_UNROLL];
        bool ok[SPLAT_UNROLL];
#pragma unroll
        for (int k = 0; k < SPLAT_UNROLL; ++k) {   // every load of the group before the first RED
            const int t = t0 + k * nworkers;
            const int ty = div_tx.quot(min(t, ntiles - 1)), tx = min(t, ntiles - 1) - ty * tiles_x;
            wi[k] = tx * 32 + lx; hi[k] = ty * 16 + ly;
            ok[k] = t < ntiles && wi[k] < W && hi[k] < H;
            fx[k] = fy[k] = 0.0f; d[k] = 1.0f;
            if (ok[k]) {
                load_flow(fs, frame, hi[k], wi[k], H, W, fx[k], fy[k]);
                if (DEPTH) d[k] = ld_stream(depth_f + (size_t)hi[k] * W + wi[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < SPLAT_UNROLL; ++k) {
            if (!ok[k]) continue;
            const Corners c = corners(wi[k], hi[k], fx[k], fy[k], W, H);
            if (!c.in_range) continue;
            const float vx = DEPTH ? -d[k] * fx[k] : -fx[k], vy = DEPTH ? -d[k] * fy[k] : -fy[k];
            atomicAdd(S + (size_t)c.T * W + c.L, make_float4(vx, vy, d[k], 0.0f));   // REDG.F32x4
        }
    }
}

// box pass + averaging of one frame, the finish worker's share: warp items of 32 columns x 8 rows (the arithmetic of
// projection_finish_kernel, same operation order).  With hole filling: bitmaps and the frame's hole list.
__device__ __forceinline__ void finish_frame(const float4 *__restrict__ Sf, float *__restrict__ cn, float *__restrict__ ou,
                                             unsigned *__restrict__ rowmask_f, unsigned char *__restrict__ colmask_f,
                                             unsigned *__restrict__ hlist, unsigned *__restrict__ h_count, unsigned frame_base,
                                             int H, int W, int gwarp, int nwarps, const FastDiv div_bw)
{
    const int lane = threadIdx.x & 31;
    const int bw = (W + 31) >> 5, nitems = bw * ((H + FIN_ROWS - 1) / FIN_ROWS);
    const size_t HW = (size_t)H * W;
    float *ov = ou + HW;
    for (int it = gwarp; it < nitems; it += nwarps) {
        const int seg = div_bw.quot(it), x = (it - seg * bw) * 32 + lane;
        const int y0 = seg * FIN_ROWS, y1 = min(y0 + FIN_ROWS, H);
        float4 c0, e0;
        load_row<true>(Sf + (size_t)max(y0 - 1, 0) * W, x, W, lane, y0 > 0, c0, e0);
        float4 prev = hsum_row(c0, e0, x, W, lane);
        unsigned colbits = 0;
#pragma unroll
        for (int ybk = 0; ybk < FIN_ROWS; ybk += FIN_UNROLL) {
            const int yb = y0 + ybk;
            float4 cur[FIN_UNROLL], edge[FIN_UNROLL];
            unsigned holes[FIN_UNROLL];
#pragma unroll
            for (int k = 0; k < FIN_UNROLL; ++k)
                load_row<true>(Sf + (size_t)min(yb + k, H - 1) * W, x, W, lane, yb + k < y1, cur[k], edge[k]);
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
                    __stcg(cn + a, sc); __stcg(ou + a, su); __stcg(ov + a, sv);
                }
                holes[k] = 0;
                if (rowmask_f) {   // uniform
                    const bool src = live && sc != 0.0f;
                    const unsigned m = __ballot_sync(0xffffffffu, src);
                    if (lane == 0 && y < y1) __stcg(rowmask_f + (size_t)y * bw + (x >> 5), m);
                    colbits |= (src ? 1u : 0u) << (y - y0);
                    holes[k] = __ballot_sync(0xffffffffu, live && !(sc > 0.0f));   // count <= 0 (:171)
                }
            }
            if (rowmask_f) {   // the group's holes go to the frame's list: one global atomic per group that has any
                unsigned total = 0;
#pragma unroll
                for (int k = 0; k < FIN_UNROLL; ++k) total += __popc(holes[k]);
                if (total) {   // uniform
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(h_count, total);
                    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
                    for (int k = 0; k < FIN_UNROLL; ++k) {
                        if (holes[k] >> lane & 1u)
                            __stcg(hlist + base + __popc(holes[k] & ((1u << lane) - 1u)), frame_base + (unsigned)((yb + k) * W + x));
                        base += __popc(holes[k]);
                    }
                }
            }
        }
        if (rowmask_f && x < W) colmask_f[(size_t)seg * W + x] = (unsigned char)colbits;   // read after the frame's barrier (cg loads)
    }
}

template <bool DEPTH>
__global__ void __launch_bounds__(NT, 2)
projection_pipeline_kernel(const FlowSource fs, const float *__restrict__ depth, float4 *__restrict__ S, int nbuf,
                           float *__restrict__ count, float *__restrict__ out, unsigned *__restrict__ rowmask,
                           unsigned char *__restrict__ colmask, unsigned *__restrict__ hlist, Ctl *__restrict__ ctl,
                           FrameCtl *__restrict__ fctl, int B, int H, int W, const FastDiv div_tx, const FastDiv div_bw)
{
    const int workers = gridDim.x >> 1;
    const bool splat_role = (int)blockIdx.x < workers;
    const int worker = splat_role ? blockIdx.x : blockIdx.x - workers;
    const size_t HW = (size_t)H * W;
    const int bw = (W + 31) >> 5, HB = (H + 7) >> 3;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (splat_role) {
        for (int f = 0; f < B; ++f) {
            if (f >= nbuf) cta_wait(&fctl[f - nbuf].c_done, workers);
            else if (f > 0) cta_wait(&ctl->pre_done, workers);
            splat_frame<DEPTH>(fs, DEPTH ? depth + (size_t)f * HW : nullptr, f, S + (size_t)(f % nbuf) * HW, H, W, worker, workers, div_tx);
            cta_signal(&fctl[f].s_done);
        }
        return;
    }
    // finish worker.  Prologue: scratch images 1 .. nbuf-1 (image 0 was cleared on the stream).
    for (size_t i = HW + (size_t)worker * NT + threadIdx.x; i < (size_t)nbuf * HW; i += (size_t)workers * NT) __stcg(S + i, zero4);
    cta_signal(&ctl->pre_done);
    const int gwarp = worker * WARPS + (threadIdx.x >> 5), nwarps = workers * WARPS;
    for (int f = 0; f < B; ++f) {
        cta_wait(&fctl[f].s_done, workers);
        finish_frame(S + (size_t)(f % nbuf) * HW, count + (size_t)f * HW, out + (size_t)f * 2 * HW,
                     rowmask ? rowmask + (size_t)f * H * bw : nullptr, colmask ? colmask + (size_t)f * HB * W : nullptr,
                     hlist, &ctl->h_count, (unsigned)((size_t)f * HW), H, W, gwarp, nwarps, div_bw);
        cta_signal(&fctl[f].f_done);
        // the image of the PREVIOUS frame goes back to the splat workers: every finish worker has long left it
        const int fc = f - 1;
        if (fc >= 0 && fc + nbuf < B) {
            cta_wait(&fctl[fc].f_done, workers);
            float4 *Sc = S + (size_t)(fc % nbuf) * HW;
            for (size_t i = (size_t)worker * NT + threadIdx.x; i < HW; i += (size_t)workers * NT) __stcg(Sc + i, zero4);
            cta_signal(&fctl[fc].c_done);
        }
    }
}

// Fills the holes the pipeline listed (entries: frame * H * W + pixel).  A separate launch: the bitmaps, counts and
// outputs are final and the scans may use the read-only path.
__global__ void __launch_bounds__(256)
projection_fill_list_kernel(const float *__restrict__ count, float *__restrict__ out, const unsigned *__restrict__ rowmask,
                            const unsigned char *__restrict__ colmask, const unsigned *__restrict__ hlist,
                            const Ctl *__restrict__ ctl, int H, int W, const FastDiv div_w)
{
    const unsigned n = ctl->h_count;
    const size_t HW = (size_t)H * W;
    const int bw = (W + 31) >> 5, HB = (H + 7) >> 3;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const unsigned p = __ldg(hlist + q);
        const unsigned fr = p / (unsigned)HW, pix = p - fr * (unsigned)HW;     // one 32-bit division per hole
        const int y = div_w.quot((int)pix), x = (int)pix - y * W;
        fill_one<false>(count + (size_t)fr * HW, out + (size_t)fr * 2 * HW, rowmask + (size_t)fr * H * bw,
                        colmask + (size_t)fr * HB * W, x, y, H, W);
    }
}

}  // namespace pipe

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
                        float *__restrict__ out, unsigned *__restrict__ rowmask, unsigned char *__restrict__ colmask, int H, int W)
{
    const int lane = threadIdx.x, x = (blockIdx.x * FIN_WARPS + threadIdx.y) * 32 + lane;
    if ((blockIdx.x * FIN_WARPS + threadIdx.y) * 32 >= W) return;   // whole warp outside
    const int b = blockIdx.z, y0 = blockIdx.y * FIN_ROWS, y1 = min(y0 + FIN_ROWS, H);
    const size_t HW = (size_t)H * W;
    const float *fu = flow + (size_t)b * 2 * HW, *fv = fu + HW;
    unsigned colbits = 0;
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
        }
    }
    if (colmask && x < W) colmask[((size_t)b * ((H + 7) >> 3) + blockIdx.y) * W + x] = (unsigned char)colbits;
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

// ---- host side of the fused pipeline ------------------------------------------------------------------------
std::atomic<int> g_projection_path{0};     // test hook: 0 = automatic, 1 = three-kernel path, 2 = fused pipeline
inline int forced_projection_path() { return g_projection_path.load(std::memory_order_relaxed); }

static int pipeline_grid(const void *kernel)
{
    int dev = 0, coop = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, pipe::NT, 0) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        return 0;
    }
    const int sms = sm_count();
    return (per_sm >= 2 ? 2 * sms : sms) & ~1;      // one splat and one finish worker per SM when both fit
}

template <bool DEPTH>
int projection_forward_pipelined(const FlowSource fs, const float *depth, float *count, float *out,
                                 int B, int H, int W, int fillhole, cudaStream_t s)
{
    using namespace pipe;
    auto kernel = projection_pipeline_kernel<DEPTH>;
    const int grid = pipeline_grid((const void *)kernel);
    const size_t HW = (size_t)H * W;
    if (grid < 2 || (fillhole && (unsigned long long)B * HW >= (1ull << 32))) return -1;   // hole-list entries are 32-bit
    const int nbuf = std::min(B, NB_MAX);
    const size_t WW = ((size_t)W + 31) >> 5, HB = ((size_t)H + 7) >> 3;
    // one stream-ordered block: [scratch images | row bitmaps | column bitmaps | hole list | counters]
    const size_t s_bytes = sizeof(float4) * HW * nbuf;
    const size_t rowmask_bytes = fillhole ? sizeof(unsigned) * B * H * WW : 0;
    const size_t colmask_bytes = fillhole ? (((size_t)B * HB * W + 15) & ~(size_t)15) : 0;
    const size_t hlist_bytes = fillhole ? sizeof(unsigned) * B * HW : 0;     // worst case: every pixel a hole
    const size_t ctl_bytes = sizeof(Ctl) + sizeof(FrameCtl) * B;
    void *mem = nullptr;
    int e = stream_scratch_alloc(&mem, s_bytes + rowmask_bytes + colmask_bytes + hlist_bytes + ctl_bytes, s);
    if (e) return e;
    char *base = static_cast<char *>(mem);
    float4 *S = reinterpret_cast<float4 *>(base);
    unsigned *rowmask = fillhole ? reinterpret_cast<unsigned *>(base + s_bytes) : nullptr;
    unsigned char *colmask = fillhole ? reinterpret_cast<unsigned char *>(base + s_bytes + rowmask_bytes) : nullptr;
    unsigned *hlist = fillhole ? reinterpret_cast<unsigned *>(base + s_bytes + rowmask_bytes + colmask_bytes) : nullptr;
    Ctl *ctl = reinterpret_cast<Ctl *>(base + s_bytes + rowmask_bytes + colmask_bytes + hlist_bytes);
    FrameCtl *fctl = reinterpret_cast<FrameCtl *>(ctl + 1);
    e = set_error(cudaMemsetAsync(S, 0, sizeof(float4) * HW, s), "clear projection scratch image 0");
    if (!e) e = set_error(cudaMemsetAsync(ctl, 0, ctl_bytes, s), "clear projection pipeline counters");
    if (!e) {
        FlowSource fsv = fs;
        const float *depth_v = depth;
        int nbuf_v = nbuf, Bv = B, Hv = H, Wv = W;
        FastDiv div_tx((unsigned)((W + 31) >> 5)), div_bw((unsigned)((W + 31) >> 5));
        void *args[] = {&fsv, &depth_v, &S, &nbuf_v, &count, &out, &rowmask, &colmask, &hlist, &ctl, &fctl, &Bv, &Hv, &Wv,
                        &div_tx, &div_bw};
        const cudaError_t le = cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(NT), args, 0, s);
        if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported) {
            (void)cudaGetLastError();
            cudaFreeAsync(mem, s);
            return -1;
        }
        e = set_error(le, "flow projection forward (pipelined)");
        if (!e) { note_launch(); e = check_launch("flow projection forward (pipelined)"); }
        if (!e && fillhole) {
            // the list length lives on the device: a fixed grid strides over it
            const unsigned nb = (unsigned)std::min<size_t>(((size_t)B * HW + 255) / 256, (size_t)sm_count() * 8);
            projection_fill_list_kernel<<<nb, 256, 0, s>>>(count, out, rowmask, colmask, hlist, ctl, H, W, FastDiv((unsigned)W));
            note_launch();
            e = check_launch("flow projection hole filling (list)");
        }
    }
    const int e2 = set_error(cudaFreeAsync(mem, s), "projection scratch (cudaFreeAsync)");
    return e ? e : e2;
}

template <bool DEPTH>
int projection_forward(const FlowSource fs, const float *depth, float *count, float *out,
                       int B, int H, int W, int fillhole, cudaStream_t s)
{
    const float *flow = fs.flow;
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !flow || !count || !out || (DEPTH && !depth)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31) || ceil_div(H, FIN_ROWS) > 65535u) return VFIDKR_ERR_ARG;
    const size_t HW = (size_t)H * W;
    if (B >= 2 && forced_projection_path() != 1) {
        const int e = projection_forward_pipelined<DEPTH>(fs, depth, count, out, B, H, W, fillhole, s);
        if (e >= 0) return e;     // -1: cooperative launch not possible here -> the three-kernel path below
    }
    // Frames are processed in chunks whose scratch image (16 B per pixel) is at most SCRATCH_BYTES; two such buffers
    // alternate and the splat of one chunk clears the buffer of the next.  Keeping the scratch L2-RESIDENT (one 1080p
    // frame, 36 MB, per chunk) was measured: the splat's DRAM traffic fell from 731 MB to its 219 MB of input, its time
    // did not move (the L2 atomic unit bounds it either way) and sixteen small launches cost 5 % more than two large
    // ones -- so the budget is generous and only very large batches are chunked.
    constexpr size_t SCRATCH_BYTES = (size_t)512 << 20;
    const int per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, SCRATCH_BYTES / (sizeof(float4) * HW)));
    const int nchunks = (B + per_chunk - 1) / per_chunk;
    const size_t chunk_cells = (size_t)per_chunk * HW;
    // one stream-ordered block: [scratch image(s) | row bitmap | column bitmap] (bitmaps only for hole filling)
    const size_t scratch_bytes = sizeof(float4) * chunk_cells * (nchunks > 1 ? 2 : 1);
    const size_t WW = ((size_t)W + 31) >> 5, HB = ((size_t)H + 7) >> 3;
    const size_t rowmask_bytes = fillhole ? sizeof(unsigned) * B * H * WW : 0, colmask_bytes = fillhole ? (size_t)B * HB * W : 0;
    void *scratch = nullptr;
    int e = stream_scratch_alloc(&scratch, scratch_bytes + rowmask_bytes + colmask_bytes, s);
    if (e) return e;
    float4 *S = static_cast<float4 *>(scratch);
    unsigned *rowmask = fillhole ? reinterpret_cast<unsigned *>(static_cast<char *>(scratch) + scratch_bytes) : nullptr;
    unsigned char *colmask = fillhole ? reinterpret_cast<unsigned char *>(rowmask) + rowmask_bytes : nullptr;
    e = set_error(cudaMemsetAsync(S, 0, sizeof(float4) * chunk_cells, s), "clear projection scratch");
    if (!e) {
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = c * per_chunk, nb = std::min(per_chunk, B - b0);
            float4 *cur = S + (size_t)(c & 1) * chunk_cells;
            float4 *nxt = (c + 1 < nchunks) ? S + (size_t)((c + 1) & 1) * chunk_cells : nullptr;
            dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), nb);
            // the next chunk may be shorter than this one; clearing nb frames of it is always enough or more
            projection_splat_kernel<DEPTH><<<grid, block, 0, s>>>(fs, b0, DEPTH ? depth + (size_t)b0 * HW : nullptr, cur, nxt, H, W);
            dim3 fblock(32, FIN_WARPS), fgrid(ceil_div(W, 32 * FIN_WARPS), ceil_div(H, FIN_ROWS), nb);
            projection_finish_kernel<<<fgrid, fblock, 0, s>>>(cur, count + (size_t)b0 * HW, out + (size_t)b0 * 2 * HW,
                                                              fillhole ? rowmask + (size_t)b0 * H * WW : nullptr,
                                                              fillhole ? colmask + (size_t)b0 * HB * W : nullptr, H, W);
            note_launch(2);
        }
        if (fillhole) {
            dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
            projection_fillhole_kernel<<<grid, block, 0, s>>>(count, out, rowmask, colmask, H, W);
            note_launch();
        }
        e = check_launch("flow projection forward");
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

VFIDKR_API int vfidkr_debug_force_projection_path(int path)
{
    if (path == 100) return pipeline_grid((const void *)pipe::projection_pipeline_kernel<true>);   // query: CTAs of the pipeline grid
    if (path < 0 || path > 2) return -1;
    return g_projection_path.exchange(path, std::memory_order_relaxed);
}

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
    const size_t HW = (size_t)H * W, WW = ((size_t)W + 31) >> 5, HB = ((size_t)H + 7) >> 3;
    const size_t key_bytes = sizeof(unsigned long long) * B * HW;
    const size_t rowmask_bytes = fillhole ? sizeof(unsigned) * B * H * WW : 0, colmask_bytes = fillhole ? (size_t)B * HB * W : 0;
    void *mem = nullptr;
    int e = stream_scratch_alloc(&mem, key_bytes + rowmask_bytes + colmask_bytes, s);
    if (e) return e;
    unsigned long long *keys = static_cast<unsigned long long *>(mem);
    unsigned *rowmask = fillhole ? reinterpret_cast<unsigned *>(static_cast<char *>(mem) + key_bytes) : nullptr;
    unsigned char *colmask = fillhole ? reinterpret_cast<unsigned char *>(rowmask) + rowmask_bytes : nullptr;
    e = set_error(cudaMemsetAsync(keys, 0, key_bytes, s), "clear min-depth keys");
    if (!e) {
        dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
        mindepth_select_kernel<<<grid, block, 0, s>>>(input1, input2, keys, H, W);
        dim3 fblock(32, FIN_WARPS), fgrid(ceil_div(W, 32 * FIN_WARPS), ceil_div(H, FIN_ROWS), B);
        mindepth_resolve_kernel<<<fgrid, fblock, 0, s>>>(keys, input1, count, output, rowmask, colmask, H, W);
        note_launch(2);
        if (fillhole) {
            projection_fillhole_kernel<<<grid, block, 0, s>>>(count, output, rowmask, colmask, H, W);
            note_launch();
        }
        e = check_launch("min-depth flow projection forward");
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
