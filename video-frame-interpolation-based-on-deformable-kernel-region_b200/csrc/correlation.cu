// correlation.cu -- FlowNet / PWC-Net cost volume, forward and backward, for sm_100a.
//
// Behaviour follows PWCNet/correlation_package_pytorch1_0/correlation_cuda_kernel.cu:47-334 and the
// shape rules of correlation_cuda.cc:23-36.  What is different:
//   * NCHW is read directly with zero padding applied on the fly -- the reference's channels_first
//     repack into padded NHWC scratch tensors (two extra passes + two fill_ kernels) is gone;
//   * fast path for the only configuration VFIDKR uses (kernel_size 1, stride1 = stride2 = 1, md = 4,
//     PWCNet.py:72): a persistent kernel with TMA-streamed shared-memory tiles of both feature maps; each
//     thread owns 4 adjacent pixels x 9 horizontal x 3 vertical displacements (108 accumulators), so a
//     float4 fetched from shared memory feeds ~11 FMAs (the reference runs 81 serial warp reductions per
//     output pixel over a padded NHWC copy);
//   * backward for the fast path keeps the 81 upstream gradients of a pixel in registers, stages the other feature
//     map in shared-memory tiles (8 channels at a time) and reuses both for every channel; one launch per gradient for
//     the whole batch (the reference launches per batch item);
//   * a generic path restates the reference formulas for any (kernel_size, stride1, stride2).
// fp32 FMA throughput, not HBM, bounds this op on SIMT (162*C flop vs 4*(2C+81) bytes per pixel);
// tcgen05 is not used: the 9-wide band of the (T+8)-wide product wastes >= 89% of a GEMM tile and
// 1e-5 parity needs a 3xTF32 split, which costs more tensor time than the SIMT kernel.
#include <algorithm>
#include <atomic>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace vfidkr {

int corr_forward_tc(const float *in1, const float *in2, float *out, float *out_b, int B, int C, int H, int W, cudaStream_t s);   // correlation_tc.cu; -1 = not applicable

namespace {

// test / measurement hook: 0 = automatic, 1 = SIMT kernels only, 2 = tensor-core kernel (tcgen05, 3 x TF32) wherever it applies
std::atomic<int> g_corr_path{0};

struct CorrShape {
    int kr, dr, ds, oc, oh, ow;
};

__host__ __device__ inline CorrShape corr_shape(int H, int W, int pad, int k, int md, int s1, int s2)
{
    CorrShape s;
    s.kr = (k - 1) / 2;
    const int border = s.kr + md;
    s.dr = md / s2;
    s.ds = 2 * s.dr + 1;
    s.oc = s.ds * s.ds;
    // ceil((padded - 2*border) / stride1) for positive numerators, as correlation_cuda.cc:31-32
    const int nh = H + 2 * pad - 2 * border, nw = W + 2 * pad - 2 * border;
    s.oh = nh > 0 ? (nh + s1 - 1) / s1 : 0;
    s.ow = nw > 0 ? (nw + s1 - 1) / s1 : 0;
    return s;
}

// zero-padded read in PADDED coordinates (what channels_first + fill_(0) provide, :47-70)
__device__ __forceinline__ float padded(const float *__restrict__ plane, int H, int W, int pad, int py, int px)
{
    const int y = py - pad, x = px - pad;
    return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(plane + (size_t)y * W + x) : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// generic forward: one thread per output pixel, loops displacements / kernel window / channels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_forward_generic_kernel(const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                            int C, int H, int W, int pad, int k, int md, int s1, int s2, CorrShape cs)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y;
    if (bx >= cs.ow || by >= cs.oh) return;
    const int n = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const float *f1 = in1 + (size_t)n * C * HW, *f2 = in2 + (size_t)n * C * HW;
    const int y1 = by * s1 + md, x1 = bx * s1 + md;   // :92-93
    const float nelems = (float)(k * k * C);          // :104
    for (int tj = -cs.dr; tj <= cs.dr; ++tj)
        for (int ti = -cs.dr; ti <= cs.dr; ++ti) {
            const int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
            float acc = 0.0f;
            for (int j = -cs.kr; j <= cs.kr; ++j)
                for (int i = -cs.kr; i <= cs.kr; ++i)
                    for (int c = 0; c < C; ++c)
                        acc += padded(f1 + (size_t)c * HW, H, W, pad, y1 + j, x1 + i) *
                               padded(f2 + (size_t)c * HW, H, W, pad, y2 + j, x2 + i);
            const int tc = (tj + cs.dr) * cs.ds + (ti + cs.dr);
            out[(((size_t)n * cs.oc + tc) * cs.oh + by) * cs.ow + bx] = acc / nelems;   // :143
        }
}

// ---------------------------------------------------------------------------------------------
// fast forward: kernel_size 1, stride1 = stride2 = 1, max_displacement 4 (9 x 9 displacements) -- the only
// configuration VFIDKR uses (PWCNet.py:72).
//
// A persistent CTA walks output tiles of TY x TX pixels and, per tile, the channel dimension in chunks of
// CK.  Each (tile, chunk) item needs an f1 tile [CK][TY][TX] and an f2 tile [CK][TY+8][TX+8]; they are
// streamed into a two-stage shared-memory ring by TMA box loads (cp.async.bulk.tensor.4d over [N][C][H][W]).
// TMA's out-of-bounds zero fill IS the op's zero padding (negative / past-the-edge coordinates, and channels
// past C), so the reference's padded NHWC repack (correlation_cuda_kernel.cu:47-70) needs no counterpart.
// Thread tile: 4 adjacent pixels x 9 horizontal x 3 vertical displacements = 108 fp32 accumulators;
// per channel it reads 1 + 3*3 float4 from shared memory (conflict-free, 16 B lane stride) for 108 FMAs.
// When W % 4 != 0 (TMA needs 16-byte row pitch) the same kernel stages the tiles with plain loads.
// ---------------------------------------------------------------------------------------------
// copies both feature maps into row-padded scratch (pitch a multiple of 4 floats) so that TMA can describe them
__global__ void __launch_bounds__(256)
corr_pad_rows_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ pa, float *__restrict__ pb,
                     size_t rows, int W, int pitch)
{
    const size_t n = rows * (size_t)pitch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / pitch;
        const int x = (int)(i - r * pitch);
        const bool in = x < W;
        pa[i] = in ? __ldg(a + r * W + x) : 0.0f;
        pb[i] = in ? __ldg(b + r * W + x) : 0.0f;
    }
}

namespace cfast {
constexpr int DR = 4, D = 2 * DR + 1;
constexpr int TX = 32, TY = 8, CK = 8, PX = 4, TJ = 3;
constexpr int F2W = TX + 2 * DR, F2H = TY + 2 * DR;
constexpr int NTHREADS = (TX / PX) * (D / TJ) * TY;          // 8 * 3 * 8 = 192
constexpr int F1_FLOATS = CK * TY * TX, F2_FLOATS = CK * F2H * F2W;
constexpr uint32_t STAGE_BYTES = (F1_FLOATS + F2_FLOATS) * sizeof(float);
// ring depth: 3 stages with two CTAs per SM for large maps; when there are fewer tiles than SMs (the coarse
// PWC levels) a CTA has the SM to itself and a 7-deep ring hides the load latency of its long channel loop
constexpr int STAGES_LARGE = 3;
constexpr size_t smem_bytes(int stages) { return (size_t)stages * STAGE_BYTES + 128; }
static_assert(D % TJ == 0 && TX % PX == 0, "tile shape");
}  // namespace cfast

template <bool TMA, int STAGES>
__global__ void __launch_bounds__(cfast::NTHREADS, STAGES <= 3 ? 2 : 1)
corr_forward_tiled_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                          const __grid_constant__ CUtensorMap map1b, const __grid_constant__ CUtensorMap map2b,
                          const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                          const float *__restrict__ in1b, const float *__restrict__ in2b, float *__restrict__ outb, int bsplit,
                          int C, int H, int W, int shift, int oh, int ow, int tiles_x, int tiles_y, int num_tiles,
                          const FastDiv div_tiles_x, const FastDiv div_tiles_image,
                          int ksplit, int cps, const FastDiv div_ksplit, float *__restrict__ partial, size_t part_stride)
{
    // Two problems of the same shape in one launch (both temporal directions of a PWC level, vfidkr_correlation_forward_pair):
    // "batch" items n >= bsplit belong to the second problem (map1b / map2b / in1b / in2b -> outb, item n - bsplit).
    // Split-K for maps with too few tiles to fill the machine: a work item is (tile, channel slice); every slice
    // walks `cps` chunks of CK channels (chunks past C are zero-filled by the loader) and, when ksplit > 1, writes
    // its UNSCALED partial sums to `partial`[slice]; corr_reduce_kernel adds the slices.  num_tiles counts
    // (tile, slice) items.  No atomics: the result does not depend on the order the slices finish in.
    using namespace cfast;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s1 = reinterpret_cast<float *>(smem_raw);                     // [STAGES][F1_FLOATS]
    float *s2 = s1 + STAGES * F1_FLOATS;                                  // [STAGES][F2_FLOATS]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s2 + STAGES * F2_FLOATS);

    const int tid = threadIdx.x;
    const int g = tid % (TX / PX);
    const int tjg = (tid / (TX / PX)) % (D / TJ);
    const int ry = tid / ((TX / PX) * (D / TJ));
    const int nchunks = cps;
    const int tiles_per_image = tiles_x * tiles_y;
    const size_t HW = (size_t)H * W;

    if (TMA && tid == 0) {
        prefetch_tensormap(&map1);
        prefetch_tensormap(&map2);
        prefetch_tensormap(&map1b);
        prefetch_tensormap(&map2b);
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto decode = [&](int vt, int &n, int &by, int &bx) {
        const int tile = div_ksplit.quot(vt);
        n = div_tiles_image.quot(tile);
        const int rem = tile - n * tiles_per_image;
        by = div_tiles_x.quot(rem);
        bx = rem - by * tiles_x;
    };
    auto issue = [&](int tile, int chunk, int stage) {   // thread 0 only (TMA path)
        int n, by, bx;
        decode(tile, n, by, bx);
        chunk += (tile - div_ksplit.quot(tile) * ksplit) * cps;   // first chunk of this item's channel slice
        const int ix0 = bx * TX + shift, iy0 = by * TY + shift;
        mbar_arrive_expect_tx(&s_full[stage], STAGE_BYTES);
        const bool second = n >= bsplit;
        tma_load_4d(s1 + stage * F1_FLOATS, second ? &map1b : &map1, &s_full[stage], ix0, iy0, chunk * CK, second ? n - bsplit : n);
        tma_load_4d(s2 + stage * F2_FLOATS, second ? &map2b : &map2, &s_full[stage], ix0 - DR, iy0 - DR, chunk * CK, second ? n - bsplit : n);
    };

    // cp.async staging of one (tile, chunk) item by all threads: one warp per tile row, lanes along x;
    // zero fill outside the image and past C.  Used when the TMA pitch requirement does not hold.
    auto stage_async = [&](int tile, int chunk, int stage) {
        int n, by, bx;
        decode(tile, n, by, bx);
        chunk += (tile - div_ksplit.quot(tile) * ksplit) * cps;   // first chunk of this item's channel slice
        const int ix0 = bx * TX + shift, iy0 = by * TY + shift;
        const int warp = tid >> 5, lane = tid & 31, nwarps = NTHREADS / 32;
        const bool second = n >= bsplit;
        const float *f1 = (second ? in1b : in1) + (size_t)(second ? n - bsplit : n) * C * HW;
        const float *f2 = (second ? in2b : in2) + (size_t)(second ? n - bsplit : n) * C * HW;
        float *d1 = s1 + stage * F1_FLOATS, *d2 = s2 + stage * F2_FLOATS;
        for (int row = warp; row < CK * TY; row += nwarps) {
            const int c = row / TY, y = row % TY, gy = iy0 + y, cc = chunk * CK + c, gx = ix0 + lane;
            const bool ok = cc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
            cp_async_4(d1 + row * TX + lane, ok ? f1 + (size_t)cc * HW + (size_t)gy * W + gx : f1, ok);
        }
        for (int row = warp; row < CK * F2H; row += nwarps) {
            const int c = row / F2H, y = row % F2H, gy = iy0 - DR + y, cc = chunk * CK + c;
            const bool rowok = cc < C && gy >= 0 && gy < H;
            const float *src = f2 + (rowok ? (size_t)cc * HW + (size_t)gy * W : 0);
            for (int x = lane; x < F2W; x += 32) {
                const int gx = ix0 - DR + x;
                const bool ok = rowok && gx >= 0 && gx < W;
                cp_async_4(d2 + row * F2W + x, ok ? src + gx : f2, ok);
            }
        }
        cp_async_commit();
    };

    // flat item stream of this CTA: item k = (tile blockIdx.x + (k / nchunks) * gridDim.x, chunk k % nchunks)
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_items = my_tiles * nchunks;
    auto produce = [&](int k) {   // start the loads of item k into stage k % STAGES (a no-op past the end)
        if (k < total_items) {
            const int t = k / nchunks, c = k - t * nchunks;
            const int tl = blockIdx.x + t * gridDim.x;
            if (TMA) { if (tid == 0) issue(tl, c, k % STAGES); }
            else stage_async(tl, c, k % STAGES);
        } else if (!TMA) {
            cp_async_commit();   // empty group keeps the wait count uniform
        }
    };
#pragma unroll 1
    for (int k = 0; k < STAGES - 1; ++k) produce(k);

    int it = 0;   // running item index (selects stage and mbarrier phase)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n, by, bx;
        decode(tile, n, by, bx);
        const int ox0 = bx * TX, oy0 = by * TY;

        // (Packed FFMA2 -- fma.rn.f32x2, 114 instead of 85.6 FMA/clk/SM in tools/microbench/ffma2.cu -- was tried here with
        // the accumulators as pairs of adjacent pixels: the b-operand pairs (columns j, j + 1) are register-aligned only for
        // even j, the odd ones cost two moves each, and the kernel ran 43 % SLOWER (PWC level 2: 278 vs 194 us).  Removed.)
        float acc[TJ][PX][D];
#pragma unroll
        for (int t = 0; t < TJ; ++t)
#pragma unroll
            for (int p = 0; p < PX; ++p)
#pragma unroll
                for (int d = 0; d < D; ++d) acc[t][p][d] = 0.0f;

        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int stage = it % STAGES;
            // keep STAGES-1 items in flight; the stage being refilled was last read in item it-1, which ended
            // with __syncthreads()
            produce(it + STAGES - 1);
            if (TMA) {
                mbar_wait(&s_full[stage], (uint32_t)((it / STAGES) & 1));
            } else {
                cp_async_wait<STAGES - 1>();   // all but the STAGES-1 most recent groups have landed
                __syncthreads();
            }
            const float *p1 = s1 + stage * F1_FLOATS + ry * TX + PX * g;
            const float *p2 = s2 + stage * F2_FLOATS + (ry + TJ * tjg) * F2W + PX * g;
#pragma unroll 2
            for (int c = 0; c < CK; ++c) {
                const float4 a4 = *reinterpret_cast<const float4 *>(p1 + c * (TY * TX));
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int t = 0; t < TJ; ++t) {
                    const float *row = p2 + (c * F2H + t) * F2W;
                    const float4 b0 = *reinterpret_cast<const float4 *>(row);
                    const float4 b1 = *reinterpret_cast<const float4 *>(row + 4);
                    const float4 b2 = *reinterpret_cast<const float4 *>(row + 8);
                    const float bv[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int p = 0; p < PX; ++p)
#pragma unroll
                        for (int d = 0; d < D; ++d) acc[t][p][d] = fmaf(av[p], bv[p + d], acc[t][p][d]);
                }
            }
            __syncthreads();   // all reads of this stage are done before it is refilled
        }

        // ---- write the 27 displacement channels of this thread: out[n, (tj, ti), oy, ox .. ox+3] ----
        const int oy = oy0 + ry, ox = ox0 + PX * g;
        if (oy < oh && ox < ow) {
            // nelems = kernel_size^2 * C (:104); the reference divides (:143), here it is one reciprocal and a
            // multiply per output (<= 1 ulp apart, far inside the 1e-5 parity bound; an IEEE division costs ~10
            // instructions x 108 outputs per thread)
            const float inv = ksplit > 1 ? 1.0f : 1.0f / (float)C;
            const size_t plane = (size_t)oh * ow;
            // direct: the problem's own output, item within the problem; split-K: the slice's partial volume, which holds
            // both problems back to back (item n)
            float *dst = n >= bsplit ? outb : out;
            int nn = n >= bsplit ? n - bsplit : n;
            if (ksplit > 1) { dst = partial + (size_t)(tile - div_ksplit.quot(tile) * ksplit) * part_stride; nn = n; }
            float *o = dst + ((size_t)nn * D * D + (size_t)(TJ * tjg) * D) * plane + (size_t)oy * ow + ox;
            const bool vec = (ow % 4 == 0) && (ox + 3 < ow) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#pragma unroll
            for (int t = 0; t < TJ; ++t)
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float *ot = o + (size_t)(t * D + d) * plane;
                    const float r[4] = {acc[t][0][d] * inv, acc[t][1][d] * inv, acc[t][2][d] * inv, acc[t][3][d] * inv};
                    if (vec) {
                        st_stream4(ot, make_float4(r[0], r[1], r[2], r[3]));
                    } else {
#pragma unroll
                        for (int p = 0; p < PX; ++p)
                            if (ox + p < ow) ot[p] = r[p];
                    }
                }
        }
    }
}

// adds the ksplit partial cost volumes and applies the 1 / (kernel_size^2 * C) scaling (:104, :143)
__global__ void __launch_bounds__(256)
corr_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out, float *__restrict__ outb, size_t n_first,
                   size_t n, int ksplit, float inv)
{
    // n elements per slice: the first n_first belong to `out`, the rest (second problem of a pair launch) to `outb`
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (int s = 0; s < ksplit; ++s) acc += __ldcs(partial + (size_t)s * n + i);
        if (i < n_first) out[i] = acc * inv;
        else outb[i - n_first] = acc * inv;
    }
}

// ---------------------------------------------------------------------------------------------
// fast backward (kernel_size 1, strides 1, max_displacement 4): one thread per input pixel keeps the 81 upstream
// gradients of that pixel in registers; the other feature map goes through shared memory in 32 x 8 tiles (+ 4 pixels
// of halo, zero padding applied by the loader) and chunks of 8 channels, so the 81 gathers per pixel and channel are LDS.
//   gi1[n,c,y,x] = 1/C * sum_tc g[n,tc,y+s,x+s]       * f2[n,c,y+tj,x+ti]          (:151-241)
//   gi2[n,c,y,x] = 1/C * sum_tc g[n,tc,y+s-tj,x+s-ti] * f1[n,c,y-tj,x-ti]          (:244-334)
// with s = pad - md, zero outside the image / the output plane.  (The first version gathered through L1 with the 81
// gradients in a local array that ptxas spilled: 34 ms for the PWC level-2 map, 0.4 % of the roofline.)
// ---------------------------------------------------------------------------------------------
namespace cbwd {
constexpr int DR = 4, D = 9, TX = 32, TY = 8, CK = 8;
constexpr int TW = TX + 2 * DR, TH = TY + 2 * DR;   // staged tile of the other map
}  // namespace cbwd

template <int WHICH>
__global__ void __launch_bounds__(cbwd::TX *cbwd::TY, 2)
corr_backward_tiled_kernel(const float *__restrict__ other, const float *__restrict__ gout, float *__restrict__ gi,
                           int C, int H, int W, int s, int oh, int ow)
{
    using namespace cbwd;
    __shared__ float tile[CK][TH][TW];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, x = x0 + tx, y = y0 + ty;
    const int n = blockIdx.z;
    const size_t HW = (size_t)H * W, plane = (size_t)oh * ow;
    const bool inside = x < W && y < H;
    const float *g = gout + (size_t)n * D * D * plane;
    float gv[D * D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) {
            const int tj = a - DR, ti = b - DR;
            const int oy = WHICH == 1 ? y + s : y + s - tj, ox = WHICH == 1 ? x + s : x + s - ti;
            gv[a * D + b] = (inside && oy >= 0 && oy < oh && ox >= 0 && ox < ow) ? __ldg(g + (size_t)(a * D + b) * plane + (size_t)oy * ow + ox) : 0.0f;
        }
    const float nel = (float)C;
    for (int c0 = 0; c0 < C; c0 += CK) {
        __syncthreads();   // the previous chunk has been consumed
        for (int i = tid; i < CK * TH * TW; i += TX * TY) {
            const int cc = i / (TH * TW), r = i - cc * (TH * TW), ry = r / TW, rx = r - ry * TW;
            const int gy = y0 - DR + ry, gx = x0 - DR + rx;
            const bool ok = c0 + cc < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
            tile[cc][ry][rx] = ok ? __ldg(other + ((size_t)n * C + c0 + cc) * HW + (size_t)gy * W + gx) : 0.0f;
        }
        __syncthreads();
        const int nc = min(CK, C - c0);
        for (int cc = 0; cc < nc; ++cc) {
            float acc = 0.0f;
#pragma unroll
            for (int a = 0; a < D; ++a)
#pragma unroll
                for (int b = 0; b < D; ++b) {
                    // gi1: f2[y + tj, x + ti] -> tile row ty + a, col tx + b;  gi2: f1[y - tj, x - ti] -> row ty + 8 - a, col tx + 8 - b
                    const float v = WHICH == 1 ? tile[cc][ty + a][tx + b] : tile[cc][ty + 2 * DR - a][tx + 2 * DR - b];
                    acc = fmaf(gv[a * D + b], v, acc);
                }
            if (inside) st_stream(gi + ((size_t)n * C + c0 + cc) * HW + (size_t)y * W + x, acc / nel);
        }
    }
}

// Register-tiled backward.  corr_backward_tiled_kernel above issues one shared-memory load per multiply-add (the 81
// upstream gradients of ONE pixel in registers, the feature taps from the tile) and is bound by the shared-memory
// pipe at ~11 % of the HBM roofline.  Here a thread owns 4 neighbouring pixels x 8 channels (32 accumulators); per
// displacement row it loads the 9 x 4 upstream gradients once (9 LDS.128, shared by its 8 channels) and a 12-float
// feature window per channel (3 LDS.128) for 36 multiply-adds: 33 LDS.128 per 288 FMAs instead of 288 LDS.32.
// A CTA owns a 32 x 8 pixel tile and stages 32 channels of the other feature map per pass (80 KB); the upstream
// gradients stream through a two-slot ring, one displacement row (9 planes) at a time, copied asynchronously while
// the previous row is being used.  For the gradient of the second input the planes are staged PRE-SHIFTED (plane d
// holds gradoutput[d] at (y - tj, x - ti), its columns starting at the 16-byte aligned x0 + 4 * floor(-ti / 4)), so
// that both gradients read aligned 16-byte windows.  The sum over the 81 displacements runs in the same order as in
// corr_backward_tiled_kernel.  256 threads, ~100 KB of shared memory: two CTAs per SM.
namespace cbr {
constexpr int DR = 4, D = 9, TX = 32, TY = 8, PX = 4, CKT = 8, NG = 4, CPASS = CKT * NG, NT = (TX / PX) * TY * NG;
constexpr int TW = TX + 2 * DR, TH = TY + 2 * DR;
constexpr int F_FLOATS = CPASS * TH * TW;
__host__ __device__ constexpr int go_pitch(int which) { return which == 1 ? TX : TX + 4; }
__host__ __device__ constexpr int go_slot(int which) { return D * TY * go_pitch(which); }   // one displacement row
constexpr size_t smem_bytes(int which) { return (size_t)(2 * go_slot(which) + F_FLOATS) * sizeof(float); }
}  // namespace cbr

// vec: bit 0 = 16-byte stores, bit 1 = 16-byte staging of the feature tile, bit 2 = 16-byte staging of gradoutput
// TMA = true (everything 16-byte aligned): one elected thread issues box loads of the feature tile and of each
// displacement row (one box of 9 planes for the first gradient, 9 shifted single-plane boxes for the second) and the
// threads wait on mbarriers; otherwise all threads stage with cp.async.
template <int WHICH, bool TMA>
__global__ void __launch_bounds__(cbr::NT, 2)
corr_backward_regtile_kernel(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CUtensorMap map_g,
                             const float *__restrict__ other, const float *__restrict__ gout, float *__restrict__ gi,
                             int C, int H, int W, int s, int oh, int ow, int vec)
{
    using namespace cbr;
    constexpr int GOP = go_pitch(WHICH), SLOT = go_slot(WHICH);
    extern __shared__ __align__(128) float smem_f[];
    __shared__ uint64_t bars[3];   // ring slots 0 / 1, feature tile
    float *sgo = smem_f;                 // [2][9][TY][GOP]
    float *sf = smem_f + 2 * SLOT;       // [CPASS][TH][TW]
    const int tid = threadIdx.x;
    const int g = tid % (TX / PX), ry = (tid / (TX / PX)) % TY, cg = tid / ((TX / PX) * TY);
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, n = blockIdx.z;
    const size_t HW = (size_t)H * W, plane = (size_t)oh * ow;
    const int vec_store = vec & 1;
    const float *gsrc = gout + (size_t)n * D * D * plane;

    // upstream gradients of the tile for displacement row a (9 planes) into ring slot a & 1; zero where there is no
    // source.  Entries of pixels outside the image are never used for a stored result.
    if (TMA) {
        if (tid == 0) {
            prefetch_tensormap(&map_f);
            prefetch_tensormap(&map_g);
            for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    auto issue_go = [&](int item, int a) {   // thread 0, TMA path: displacement row a into ring slot item & 1
        uint64_t *bar = &bars[item & 1];
        float *dst = sgo + (item & 1) * SLOT;
        mbar_arrive_expect_tx(bar, SLOT * sizeof(float));
        if (WHICH == 1) {
            tma_load_4d(dst, &map_g, bar, x0 + s, y0 + s, a * D, n);
        } else {
#pragma unroll
            for (int bb = 0; bb < D; ++bb)
                tma_load_4d(dst + bb * TY * GOP, &map_g, bar, x0 + s + 4 - 4 * ((bb + 3) / 4), y0 + s - (a - DR), a * D + bb, n);
        }
    };
    auto stage_go = [&](int item, int a) {
        float *dst = sgo + (item & 1) * SLOT;
        const int tj = a - DR;
        if (vec & 4) {   // 16-byte groups: W, ow and s are multiples of 4, so a group is all in or all out
            for (int i = tid; i < D * TY * (GOP / 4); i += NT) {
                const int bb = i / (TY * (GOP / 4)), r = i - bb * (TY * (GOP / 4)), sy = r / (GOP / 4), sx = 4 * (r - sy * (GOP / 4));
                const int oy = WHICH == 1 ? y0 + sy + s : y0 + sy + s - tj;
                const int ox = WHICH == 1 ? x0 + sx + s : x0 + sx + s + 4 - 4 * ((bb + 3) / 4);   // 4 * floor((4 - b) / 4) = 4, 0, -4
                const bool ok = oy >= 0 && oy < oh && ox >= 0 && ox < ow;
                cp_async_16(dst + (bb * TY + sy) * GOP + sx, ok ? gsrc + (size_t)(a * D + bb) * plane + (size_t)oy * ow + ox : gsrc, ok);
            }
        } else {
            for (int i = tid; i < SLOT; i += NT) {
                const int bb = i / (TY * GOP), r = i - bb * (TY * GOP), sy = r / GOP, sx = r - sy * GOP;
                const int oy = WHICH == 1 ? y0 + sy + s : y0 + sy + s - tj;
                const int ox = WHICH == 1 ? x0 + sx + s : x0 + sx + s + 4 - 4 * ((bb + 3) / 4);
                const bool ok = oy >= 0 && oy < oh && ox >= 0 && ox < ow;
                cp_async_4(dst + i, ok ? gsrc + (size_t)(a * D + bb) * plane + (size_t)oy * ow + ox : gsrc, ok);
            }
        }
    };

    // sum / nelems as the reference writes it; for a power-of-two channel count the reciprocal multiply is bit-identical
    const float nel = (float)C, rnel = 1.0f / nel;
    const bool pow2 = (C & (C - 1)) == 0;
    auto scaled = [&](float v) { return pow2 ? v * rnel : v / nel; };
    int item = 0, pass = 0;   // displacement rows staged so far (ring position / mbarrier phase), channel passes
    for (int c0 = 0; c0 < C; c0 += CPASS, ++pass) {
        // (the previous pass ended with a barrier: the feature tile and both ring slots are free)
        if (TMA) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&bars[2], F_FLOATS * sizeof(float));
                tma_load_4d(sf, &map_f, &bars[2], x0 - DR, y0 - DR, c0, n);   // rows / columns / channels outside: zero
                issue_go(item, 0);
            }
        } else if (vec & 2) {   // W % 4 == 0 and x0 - DR a multiple of 4: a group of four is all inside the image or all outside
            for (int i = tid; i < CPASS * TH * (TW / 4); i += NT) {
                const int ch = i / (TH * (TW / 4)), r2 = i - ch * (TH * (TW / 4)), r = r2 / (TW / 4), cx = 4 * (r2 - r * (TW / 4));
                const int gy = y0 - DR + r, gx = x0 - DR + cx;
                const bool ok = c0 + ch < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
                cp_async_16(sf + (ch * TH + r) * TW + cx, ok ? other + ((size_t)n * C + c0 + ch) * HW + (size_t)gy * W + gx : other, ok);
            }
        } else {
            for (int ch = 0; ch < CPASS; ++ch) {
                const bool chok = c0 + ch < C;
                const float *src = other + ((size_t)n * C + (chok ? c0 + ch : 0)) * HW;
                for (int i = tid; i < TH * TW; i += NT) {
                    const int r = i / TW, cx = i - r * TW, gy = y0 - DR + r, gx = x0 - DR + cx;
                    const bool ok = chok && gy >= 0 && gy < H && gx >= 0 && gx < W;
                    cp_async_4(sf + ch * (TH * TW) + i, ok ? src + (size_t)gy * W + gx : other, ok);
                }
            }
        }
        if (!TMA) {
            stage_go(item, 0);
            cp_async_commit();
        }

        float acc[CKT][PX];
#pragma unroll
        for (int ch = 0; ch < CKT; ++ch)
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[ch][p] = 0.0f;
#pragma unroll 1
        for (int a = 0; a < D; ++a, ++item) {
            // the next row goes into the slot last read in iteration a - 1, which ended with a barrier
            if (TMA) {
                if (tid == 0 && a + 1 < D) issue_go(item + 1, a + 1);
                if (a == 0) mbar_wait(&bars[2], (uint32_t)(pass & 1));
                mbar_wait(&bars[item & 1], (uint32_t)((item >> 1) & 1));
            } else {
                if (a + 1 < D) {
                    stage_go(item + 1, a + 1);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
            }
            const float *go_a = sgo + (item & 1) * SLOT;
            float gq[D][PX];
#pragma unroll
            for (int b = 0; b < D; ++b) {
                const float *gp = go_a + (b * TY + ry) * GOP + PX * g;
                const float4 t = *reinterpret_cast<const float4 *>(gp);
                if (WHICH == 1 || (4 - b) % 4 == 0) {
                    gq[b][0] = t.x; gq[b][1] = t.y; gq[b][2] = t.z; gq[b][3] = t.w;
                } else {   // window at column offset o = (4 - b) mod 4 of the aligned plane: two aligned loads
                    const float4 u = *reinterpret_cast<const float4 *>(gp + 4);
                    const float w8[8] = {t.x, t.y, t.z, t.w, u.x, u.y, u.z, u.w};
                    const int o = ((4 - b) % 4 + 4) % 4;
#pragma unroll
                    for (int p = 0; p < PX; ++p) gq[b][p] = w8[o + p];
                }
            }
            // gi1: f2[y + tj, x + ti] -> tile row ry + a, column px + b;  gi2: f1[y - tj, x - ti] -> row ry + 8 - a, column px + 8 - b
            const float *frow = sf + (cg * CKT) * (TH * TW) + (WHICH == 1 ? ry + a : ry + 2 * DR - a) * TW + PX * g;
#pragma unroll
            for (int ch = 0; ch < CKT; ++ch) {
                const float4 f0 = *reinterpret_cast<const float4 *>(frow + ch * (TH * TW));
                const float4 f1 = *reinterpret_cast<const float4 *>(frow + ch * (TH * TW) + 4);
                const float4 f2 = *reinterpret_cast<const float4 *>(frow + ch * (TH * TW) + 8);
                const float fv[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
#pragma unroll
                for (int b = 0; b < D; ++b)
#pragma unroll
                    for (int p = 0; p < PX; ++p)
                        acc[ch][p] = fmaf(gq[b][p], fv[WHICH == 1 ? p + b : p + 2 * DR - b], acc[ch][p]);
            }
            __syncthreads();   // this slot (and, after the last row, the feature tile) may be overwritten
        }
        const int y = y0 + ry, x = x0 + PX * g;
        if (y < H) {
#pragma unroll
            for (int ch = 0; ch < CKT; ++ch) {
                const int c = c0 + cg * CKT + ch;
                if (c >= C) break;
                float *dst = gi + ((size_t)n * C + c) * HW + (size_t)y * W + x;
                if (vec_store && x + PX <= W) {
                    st_stream4(dst, make_float4(scaled(acc[ch][0]), scaled(acc[ch][1]), scaled(acc[ch][2]), scaled(acc[ch][3])));
                } else {
#pragma unroll
                    for (int p = 0; p < PX; ++p)
                        if (x + p < W) st_stream(dst + p, scaled(acc[ch][p]));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// generic backward: restates the reference formulas, one thread per (n, c, y, x); stride1 must be 1
// for the reference itself to stay in bounds, other strides follow the same index arithmetic with
// out-of-plane targets skipped.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_backward_generic_kernel(const float *__restrict__ in1, const float *__restrict__ in2,
                             const float *__restrict__ gout, float *__restrict__ gi1, float *__restrict__ gi2,
                             int C, int H, int W, int pad, int k, int md, int s1, int s2, CorrShape cs)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y;
    if (bx >= W || by >= H) return;
    const int n = blockIdx.z / C, c = blockIdx.z % C;
    const size_t HW = (size_t)H * W, plane = (size_t)cs.oh * cs.ow;
    const float *p1 = in1 + ((size_t)n * C + c) * HW, *p2 = in2 + ((size_t)n * C + c) * HW;
    const float *g = gout + (size_t)n * cs.oc * plane;
    const int y = by * s1 + pad, x = bx * s1 + pad;   // :162-163
    const float nelems = (float)(k * k * C);
    const int oy = y - pad, ox = x - pad;
    const bool target_ok = oy >= 0 && oy < H && ox >= 0 && ox < W;
    float sum1 = 0.0f, sum2 = 0.0f;
    {   // input1 (:170-236); C integer division truncates toward zero
        int xmin = (x - cs.kr - md) / s1, ymin = (y - cs.kr - md) / s1;
        int xmax = (x + cs.kr - md) / s1, ymax = (y + cs.kr - md) / s1;
        if (!(xmax < 0 || ymax < 0 || xmin >= cs.ow || ymin >= cs.oh) && !(xmin > xmax || ymin > ymax)) {
            xmin = max(0, xmin); xmax = min(cs.ow - 1, xmax);
            ymin = max(0, ymin); ymax = min(cs.oh - 1, ymax);
            for (int tc = 0; tc < cs.oc; ++tc) {
                const int i2 = (tc % cs.ds - cs.dr) * s2, j2 = (tc / cs.ds - cs.dr) * s2;
                const float val2 = padded(p2, H, W, pad, y + j2, x + i2);
                for (int j = ymin; j <= ymax; ++j)
                    for (int i = xmin; i <= xmax; ++i) sum1 += __ldg(g + (size_t)tc * plane + (size_t)j * cs.ow + i) * val2;
            }
        }
    }
    for (int tc = 0; tc < cs.oc; ++tc) {   // input2 (:285-320)
        const int i2 = (tc % cs.ds - cs.dr) * s2, j2 = (tc / cs.ds - cs.dr) * s2;
        int xmin = (x - cs.kr - md - i2) / s1, ymin = (y - cs.kr - md - j2) / s1;
        int xmax = (x + cs.kr - md - i2) / s1, ymax = (y + cs.kr - md - j2) / s1;
        if (xmax < 0 || ymax < 0 || xmin >= cs.ow || ymin >= cs.oh) continue;
        if (xmin > xmax || ymin > ymax) continue;
        xmin = max(0, xmin); xmax = min(cs.ow - 1, xmax);
        ymin = max(0, ymin); ymax = min(cs.oh - 1, ymax);
        const float val1 = padded(p1, H, W, pad, y - j2, x - i2);
        for (int j = ymin; j <= ymax; ++j)
            for (int i = xmin; i <= xmax; ++i) sum2 += __ldg(g + (size_t)tc * plane + (size_t)j * cs.ow + i) * val1;
    }
    if (target_ok) {
        gi1[((size_t)n * C + c) * HW + (size_t)oy * W + ox] = sum1 / nelems;
        gi2[((size_t)n * C + c) * HW + (size_t)oy * W + ox] = sum2 / nelems;
    }
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_correlation_outshape(int H, int W, int pad, int k, int md, int s1, int s2,
                                           int *oc, int *oh, int *ow)
{
    if (H <= 0 || W <= 0 || pad < 0 || k <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || !oc || !oh || !ow) return VFIDKR_ERR_ARG;
    const CorrShape cs = corr_shape(H, W, pad, k, md, s1, s2);
    *oc = cs.oc; *oh = cs.oh; *ow = cs.ow;
    return VFIDKR_OK;
}

// fast path (kernel_size 1, strides 1, max_displacement 4): output = corr(input1, input2) and, when output_b is given,
// output_b = corr(input2, input1) in the SAME launch (twice the tiles: the coarse PWC levels fill the machine, no
// second launch / second split-K reduce)
static int corr_forward_fast(const float *input1, const float *input2, float *output, float *output_b,
                             int B, int C, int H, int W, int pad, int md, const CorrShape &cs, cudaStream_t s)
{
    using namespace cfast;
    const int nprob = output_b ? 2 : 1;
    const int tiles_x = ceil_div(cs.ow, TX), tiles_y = ceil_div(cs.oh, TY);
    const long long num_tiles = (long long)tiles_x * tiles_y * B * nprob;
    if (num_tiles >= (1ll << 31)) return VFIDKR_ERR_ARG;
    const FastDiv dx((unsigned)tiles_x), di((unsigned)(tiles_x * tiles_y));
    CUtensorMap m1, m2, m1b, m2b;
    // (TMA box origins are kept at multiples of 4 elements: pad != max_displacement shifts them by md - pad, and an
    // origin of -2 was observed to fault -- such configurations, which VFIDKR never uses, take the cp.async path)
    const bool tma_ok = (md - pad) % 4 == 0;
    bool tma = tma_ok && (W % 4 == 0) && aligned16(input1) && aligned16(input2) &&
               encode_tensor_map_4d(&m1, input1, W, H, C, B, TX, TY, CK) &&
               encode_tensor_map_4d(&m2, input2, W, H, C, B, F2W, F2H, CK) &&
               (!output_b || (encode_tensor_map_4d(&m1b, input2, W, H, C, B, TX, TY, CK) &&
                              encode_tensor_map_4d(&m2b, input1, W, H, C, B, F2W, F2H, CK)));
    // Rows that are not a multiple of 16 bytes (the two coarsest PWC levels at 1080p are 62 and 31 wide) cannot be
    // described to TMA.  Both feature maps are then copied once into a row-padded scratch (one small kernel, a few
    // MB) and the tensor maps keep the true width as extent -- reads past it are zero-filled, which is the op's own
    // padding -- so these levels run the TMA kernel too instead of the 4-byte cp.async staging path.
    void *pitched = nullptr;
    // (worth it between ~1.5 M and 64 M elements per map: below, the extra launch costs more than the staging path
    // loses -- measured 44 -> 50 us at 18 x 31 x 196 x 8, 66 -> 51 us at 36 x 62 x 128 x 8 -- above, the copy is real
    // traffic; a pair launch reads each padded map twice, so it pays from half that size)
    const size_t map_elems = (size_t)B * C * H * W;
    if (!tma && tma_ok && map_elems * nprob >= ((size_t)3 << 19) && map_elems <= ((size_t)64 << 20)) {
        const size_t pitch = ((size_t)W + 3) & ~(size_t)3, rows = (size_t)B * C * H;
        if (stream_scratch_alloc(&pitched, 2 * rows * pitch * sizeof(float), s) == VFIDKR_OK) {
            float *p1 = static_cast<float *>(pitched), *p2 = p1 + rows * pitch;
            const unsigned nb = (unsigned)std::min<size_t>((rows * pitch + 255) / 256, (size_t)sm_count() * 16);
            corr_pad_rows_kernel<<<nb, 256, 0, s>>>(input1, input2, p1, p2, rows, W, (int)pitch);
            note_launch();
            tma = cudaGetLastError() == cudaSuccess &&
                  encode_tensor_map_4d_pitched(&m1, p1, W, H, C, B, pitch, TX, TY, CK) &&
                  encode_tensor_map_4d_pitched(&m2, p2, W, H, C, B, pitch, F2W, F2H, CK) &&
                  (!output_b || (encode_tensor_map_4d_pitched(&m1b, p2, W, H, C, B, pitch, TX, TY, CK) &&
                                 encode_tensor_map_4d_pitched(&m2b, p1, W, H, C, B, pitch, F2W, F2H, CK)));
        }
        (void)cudaGetLastError();
    }
    struct PitchedGuard {   // the scratch is released on the stream, after the kernel that reads it
        void *p; cudaStream_t s;
        ~PitchedGuard() { if (p) cudaFreeAsync(p, s); }
    } pitched_guard{pitched, s};
    if (!tma) { memset(&m1, 0, sizeof m1); memset(&m2, 0, sizeof m2); }
    if (!tma || !output_b) { m1b = m1; m2b = m2; }
    // Too few tiles to fill the machine (the two coarsest PWC levels at 1080p): split the channels over
    // ksplit work items per tile and add the partial volumes in a second, tiny kernel.
    const int nchunks = (C + CK - 1) / CK;
    // The slice count is chosen by cost, not by "as many items as CTA slots": an item costs its chunks plus ~2
    // chunk-times of pipeline fill / partial write, and items are dealt in rounds of (2 x SMs).  (80 tiles x 16
    // chunks at PWC level 5: 4 slices made 320 items = 2 rounds of 4 chunks; 3 slices make 240 items = 1 round of 6.)
    int ksplit = 1;
    if (num_tiles <= (long long)sm_count()) {
        const long long slots = 2ll * sm_count();
        long long best = -1;
        for (int ks = 1; ks <= nchunks; ++ks) {
            const int cp = (nchunks + ks - 1) / ks, eff = (nchunks + cp - 1) / cp;
            const long long rounds = (num_tiles * eff + slots - 1) / slots;
            const long long cost = rounds * (cp + 2) + (eff > 1 ? 1 : 0);   // + the reduce kernel
            if (best < 0 || cost < best) { best = cost; ksplit = eff; }
        }
    }
    const int cps = (nchunks + ksplit - 1) / ksplit;
    ksplit = (nchunks + cps - 1) / cps;   // drop slices that would be empty
    const long long num_items = num_tiles * ksplit;
    const size_t out_elems = (size_t)B * D * D * cs.oh * cs.ow;     // one problem
    void *partial = nullptr;
    if (ksplit > 1) {
        const int e = stream_scratch_alloc(&partial, sizeof(float) * out_elems * nprob * ksplit, s);
        if (e) return e;
    }
    auto launch = [&](auto kernel, int stages, int ctas_per_sm) {
        // raising the dynamic shared memory limit is idempotent and cheap; do it on every launch so the
        // attribute is set on whichever device is current
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(stages));
        const int nblk = (int)std::min<long long>(num_items, (long long)sm_count() * ctas_per_sm);
        kernel<<<nblk, NTHREADS, smem_bytes(stages), s>>>(m1, m2, m1b, m2b, input1, input2, output, input2, input1, output_b, B,
                                                          C, H, W, md - pad, cs.oh, cs.ow,
                                                          tiles_x, tiles_y, (int)num_items, dx, di, ksplit, cps,
                                                          FastDiv((unsigned)ksplit), static_cast<float *>(partial), out_elems * nprob);
    };
    if (tma) launch(corr_forward_tiled_kernel<true, STAGES_LARGE>, STAGES_LARGE, 2);
    else     launch(corr_forward_tiled_kernel<false, STAGES_LARGE>, STAGES_LARGE, 2);
    note_launch();
    if (ksplit > 1) {
        int e = check_launch("correlation forward (split-K)");
        if (!e) {
            const size_t n_all = out_elems * nprob;
            const unsigned nb = (unsigned)std::min<size_t>((n_all + 255) / 256, (size_t)sm_count() * 8);
            corr_reduce_kernel<<<nb, 256, 0, s>>>(static_cast<const float *>(partial), output, output_b, out_elems, n_all, ksplit, 1.0f / (float)C);
            note_launch();
            e = check_launch("correlation forward (reduce)");
        }
        const int e2 = set_error(cudaFreeAsync(partial, s), "free correlation scratch");
        return e ? e : e2;
    }
    return check_launch("correlation forward");
}

// Which kernel serves a (pad = md = 4, k = 1) forward.  Measured on B200 (profiles/r02/time_corr_tensor_v3.log, B = 8):
// the tcgen05 kernel wins where the maps are small and deep -- PWC levels 6 and 5, 196 x 18 x 31 and 128 x 36 x 62:
// 34.8 vs 48.4 us and 37.0 vs 55.1 us, the FFMA kernel needing split-K plus a reduction there -- and loses from level 4
// on (46 vs 40, 123 vs 95, 284 vs 198 us), where its operand staging (global -> registers -> hi / lo split ->
// shared memory) and not the tensor pipe sets the pace.  path 1 / 2 (vfidkr_debug_force_correlation_path) force one.
static bool use_tensor_path(int B, int C, int H, int W)
{
    const int path = g_corr_path.load(std::memory_order_relaxed);
    if (path) return path == 2;
    return C >= 96 && (long long)B * H * W <= 40000;
}

VFIDKR_API int vfidkr_correlation_forward(const float *input1, const float *input2, float *output,
                                          int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                          int corr_type_multiply, vfidkr_stream_t stream)
{
    (void)corr_type_multiply;   // accepted and ignored, as in the reference kernels
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || k <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !output) return VFIDKR_ERR_ARG;
    const CorrShape cs = corr_shape(H, W, pad, k, md, s1, s2);
    if (cs.oh <= 0 || cs.ow <= 0) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4 && pad == 4 && use_tensor_path(B, C, H, W)) {
        const int e = corr_forward_tc(input1, input2, output, nullptr, B, C, H, W, s);
        if (e >= 0) return e;
    }
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4) return corr_forward_fast(input1, input2, output, nullptr, B, C, H, W, pad, md, cs, s);
    dim3 block(32, 8), grid(ceil_div(cs.ow, 32), ceil_div(cs.oh, 8), B);
    corr_forward_generic_kernel<<<grid, block, 0, s>>>(input1, input2, output, C, H, W, pad, k, md, s1, s2, cs);
    note_launch();
    return check_launch("correlation forward");
}

VFIDKR_API int vfidkr_debug_force_correlation_path(int path)
{
    if (path < 0 || path > 2) return -1;
    return g_corr_path.exchange(path, std::memory_order_relaxed);
}

// Both temporal directions of a pyramid level at once: output12 = correlation(input1, input2), output21 =
// correlation(input2, input1) (PWC-Net is run once per direction, networks/DAIN.py:196-202, and calls the correlation
// with the roles of the two feature maps exchanged; when the second map is not warped -- pyramid level 6,
// PWCNet/PWCNet.py:230 -- or a caller evaluates both directions level by level, the two calls share their inputs).
VFIDKR_API int vfidkr_correlation_forward_pair(const float *input1, const float *input2, float *output12, float *output21,
                                               int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                               int corr_type_multiply, vfidkr_stream_t stream)
{
    (void)corr_type_multiply;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || k <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || 2 * B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !output12 || !output21) return VFIDKR_ERR_ARG;
    const CorrShape cs = corr_shape(H, W, pad, k, md, s1, s2);
    if (cs.oh <= 0 || cs.ow <= 0) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4 && pad == 4 && use_tensor_path(2 * B, C, H, W)) {
        const int e = corr_forward_tc(input1, input2, output12, output21, B, C, H, W, s);
        if (e >= 0) return e;
    }
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4) return corr_forward_fast(input1, input2, output12, output21, B, C, H, W, pad, md, cs, s);
    dim3 block(32, 8), grid(ceil_div(cs.ow, 32), ceil_div(cs.oh, 8), B);
    corr_forward_generic_kernel<<<grid, block, 0, s>>>(input1, input2, output12, C, H, W, pad, k, md, s1, s2, cs);
    corr_forward_generic_kernel<<<grid, block, 0, s>>>(input2, input1, output21, C, H, W, pad, k, md, s1, s2, cs);
    note_launch(2);
    return check_launch("correlation forward (pair)");
}

VFIDKR_API int vfidkr_correlation_backward(const float *input1, const float *input2, const float *gradoutput,
                                           float *gradinput1, float *gradinput2,
                                           int B, int C, int H, int W, int pad, int k, int md, int s1, int s2,
                                           int corr_type_multiply, vfidkr_stream_t stream)
{
    (void)corr_type_multiply;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || pad < 0 || k <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !gradoutput || !gradinput1 || !gradinput2) return VFIDKR_ERR_ARG;
    const CorrShape cs = corr_shape(H, W, pad, k, md, s1, s2);
    if (cs.oh <= 0 || cs.ow <= 0) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4 && pad >= md) {
        dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B), grid_r(ceil_div(W, cbr::TX), ceil_div(H, cbr::TY), B);
        const bool w4 = W % 4 == 0;
        const int vgo = (w4 && cs.ow % 4 == 0 && (pad - md) % 4 == 0 && aligned16(gradoutput)) ? 4 : 0;
        const int v1 = ((w4 && aligned16(gradinput1)) ? 1 : 0) | ((w4 && aligned16(input2)) ? 2 : 0) | vgo;
        const int v2 = ((w4 && aligned16(gradinput2)) ? 1 : 0) | ((w4 && aligned16(input1)) ? 2 : 0) | vgo;
        CUtensorMap mf1, mf2, mg1, mg2;
        memset(&mf1, 0, sizeof(mf1)); memset(&mf2, 0, sizeof(mf2)); memset(&mg1, 0, sizeof(mg1)); memset(&mg2, 0, sizeof(mg2));
        const bool tma1 = (v1 & 6) == 6 && encode_tensor_map_4d(&mf1, input2, W, H, C, B, cbr::TW, cbr::TH, cbr::CPASS) &&
                          encode_tensor_map_4d(&mg1, gradoutput, cs.ow, cs.oh, 81, B, cbr::go_pitch(1), cbr::TY, cbr::D);
        const bool tma2 = (v2 & 6) == 6 && encode_tensor_map_4d(&mf2, input1, W, H, C, B, cbr::TW, cbr::TH, cbr::CPASS) &&
                          encode_tensor_map_4d(&mg2, gradoutput, cs.ow, cs.oh, 81, B, cbr::go_pitch(2), cbr::TY, 1);
        auto k1 = tma1 ? corr_backward_regtile_kernel<1, true> : corr_backward_regtile_kernel<1, false>;
        auto k2 = tma2 ? corr_backward_regtile_kernel<2, true> : corr_backward_regtile_kernel<2, false>;
        const bool big_smem =
            cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbr::smem_bytes(1)) == cudaSuccess &&
            cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbr::smem_bytes(2)) == cudaSuccess;
        if (big_smem && ceil_div(H, cbr::TY) <= 65535u) {
            k1<<<grid_r, cbr::NT, cbr::smem_bytes(1), s>>>(mf1, mg1, input2, gradoutput, gradinput1, C, H, W, pad - md, cs.oh, cs.ow, v1);
            k2<<<grid_r, cbr::NT, cbr::smem_bytes(2), s>>>(mf2, mg2, input1, gradoutput, gradinput2, C, H, W, pad - md, cs.oh, cs.ow, v2);
        } else {
            (void)cudaGetLastError();
            corr_backward_tiled_kernel<1><<<grid, block, 0, s>>>(input2, gradoutput, gradinput1, C, H, W, pad - md, cs.oh, cs.ow);
            corr_backward_tiled_kernel<2><<<grid, block, 0, s>>>(input1, gradoutput, gradinput2, C, H, W, pad - md, cs.oh, cs.ow);
        }
        note_launch(2);
    } else {
        if ((long long)B * C > 65535) return VFIDKR_ERR_ARG;
        // targets outside the plane are skipped, so clear first (the reference fill_(0)s, correlation_cuda.cc:111-113)
        int e = set_error(cudaMemsetAsync(gradinput1, 0, sizeof(float) * (size_t)B * C * H * W, s), "clear gradinput1");
        if (e) return e;
        e = set_error(cudaMemsetAsync(gradinput2, 0, sizeof(float) * (size_t)B * C * H * W, s), "clear gradinput2");
        if (e) return e;
        dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B * C);
        corr_backward_generic_kernel<<<grid, block, 0, s>>>(input1, input2, gradoutput, gradinput1, gradinput2, C, H, W, pad, k, md, s1, s2, cs);
        note_launch();
    }
    return check_launch("correlation backward");
}
