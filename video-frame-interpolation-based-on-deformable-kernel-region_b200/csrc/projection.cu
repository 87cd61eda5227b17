// projection.cu -- FlowProjection / DepthFlowProjection (forward splat with atomics, averaging,
// hole filling; gather backward) for sm_100a.
//
// Behaviour follows my_package/FlowProjection/flowprojection_cuda_kernel.cu:29-301 and
// my_package/DepthFlowProjection/depthflowprojection_cuda_kernel.cu:29-341.  One templated kernel
// family serves both (DEPTH = false -> weight 1).  Atomics are kept exactly where the reference
// splats (the four corners of (x+fx, y+fy), three planes each); what changes:
//   * the 2 x 2 splat is split into an atomic vertical half (rows T, Bm at column L: 6 REDs per pixel
//     instead of 12) and a dense, atomic-free horizontal half fused with the averaging pass -- same sums,
//     half the atomics, which is what bounds this op (the L2 retires ~1 fp32 RED per clock per slice);
//   * the accumulation planes are cleared by the library on the stream (no caller zero-fill);
//   * the backward is a pure gather with register accumulation and a single store per output
//     (the reference does eight / sixteen read-modify-writes of its own pixel).
#include <algorithm>

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
// cells {T, Bm} x {L, R} (:75-88).  The L2 atomic units retire about one fp32 RED per clock per slice, so the
// splat is bound by the NUMBER of atomics (12 per pixel in the reference).  Here only the left column of the
// block is splatted -- rows T and Bm at column L, 6 atomics per pixel, 3 when the border clamp makes Bm == T --
// and the right column is produced by the dense pass below:  A[y][x] = S[y][x] + S[y][x-1],
// with the reference's border behaviour (R = min(L+1, W-1), so a pixel with L == W-1 hits column W-1 twice):
//   A[y][W-1] = 2*S[y][W-1] + S[y][W-2].
template <bool DEPTH>
__global__ void __launch_bounds__(BX *BY)
projection_splat_kernel(const float *__restrict__ flow, const float *__restrict__ depth,
                        float *__restrict__ count, float *__restrict__ out, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, pix = (size_t)h_i * W + w_i;
    const float fx = ld_stream(flow + ((size_t)b * 2 + 0) * HW + pix);
    const float fy = ld_stream(flow + ((size_t)b * 2 + 1) * HW + pix);
    const Corners c = corners(w_i, h_i, fx, fy, W, H);
    if (!c.in_range) return;
    const float d = DEPTH ? ld_stream(depth + (size_t)b * HW + pix) : 1.0f;
    const float vx = DEPTH ? -d * fx : -fx, vy = DEPTH ? -d * fy : -fy;   // :75-88 / depth :77-92
    float *ou = out + ((size_t)b * 2 + 0) * HW, *ov = ou + HW, *cn = count + (size_t)b * HW;
    const size_t a0 = (size_t)c.T * W + c.L;
    if (c.Bm == c.T) {   // bottom row clamped onto the top row: the cell is hit twice
        red_add(ou + a0, 2.0f * vx);
        red_add(ov + a0, 2.0f * vy);
        red_add(cn + a0, 2.0f * d);
    } else {
        const size_t a1 = a0 + W;
        red_add(ou + a0, vx); red_add(ou + a1, vx);
        red_add(ov + a0, vy); red_add(ov + a1, vy);
        red_add(cn + a0, d);  red_add(cn + a1, d);
    }
}

// Dense pass: completes the horizontal half of the 2 x 2 splat (see above) and averages (:128-135), in place.
// One CTA owns whole rows and walks each row right to left in segments, so a cell is always read before the
// thread that rewrites it runs: no scratch buffer is needed and there is no dependence between CTAs.
constexpr int ROW_THREADS = 256;

__global__ void __launch_bounds__(ROW_THREADS)
projection_finish_kernel(float *__restrict__ count, float *__restrict__ out, int H, int W, int rows_total)
{
    const size_t HW = (size_t)H * W;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const int b = row / H, y = row - b * H;
        float *ou = out + ((size_t)b * 2 + 0) * HW + (size_t)y * W, *ov = ou + HW;
        float *cn = count + (size_t)b * HW + (size_t)y * W;
        const int nseg = (W + ROW_THREADS - 1) / ROW_THREADS;
        for (int seg = nseg - 1; seg >= 0; --seg) {
            const int x = seg * ROW_THREADS + (int)threadIdx.x;
            float su = 0.f, sv = 0.f, sc = 0.f;
            if (x < W) {
                const float w0 = (x == W - 1) ? 2.0f : 1.0f;   // clamped right corner lands on column W-1 again
                su = w0 * ou[x]; sv = w0 * ov[x]; sc = w0 * cn[x];
                if (x > 0) { su += ou[x - 1]; sv += ov[x - 1]; sc += cn[x - 1]; }
            }
            __syncthreads();   // every read of this segment (and of the cell left of it) precedes the writes
            if (x < W) {
                cn[x] = sc;
                if (sc > 0.0f) { su = su / sc; sv = sv / sc; }   // :130-134
                ou[x] = su; ov[x] = sv;
            }
            // the next segment (to the left) only reads cells this segment did not write
        }
    }
}

// hole filling (:171-232).  Reads only non-hole pixels (count != 0), which this kernel never writes,
// so running it in place is race-free.  The four scans are unbounded as in the reference.
__global__ void __launch_bounds__(BX *BY)
projection_fillhole_kernel(const float *__restrict__ count, float *__restrict__ out, int H, int W)
{
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= W || h_i >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const float *cn = count + (size_t)b * HW;
    if (cn[(size_t)h_i * W + w_i] > 0.0f) return;
    int lo = w_i; float lt = 0.0f;
    while (lt == 0.0f && lo - 1 >= 0) { --lo; lt = cn[(size_t)h_i * W + lo]; }
    int ro = w_i; float rt = 0.0f;
    while (rt == 0.0f && ro + 1 <= W - 1) { ++ro; rt = cn[(size_t)h_i * W + ro]; }
    int uo = h_i; float ut = 0.0f;
    while (ut == 0.0f && uo - 1 >= 0) { --uo; ut = cn[(size_t)uo * W + w_i]; }
    int dn = h_i; float dt = 0.0f;
    while (dt == 0.0f && dn + 1 <= H - 1) { ++dn; dt = cn[(size_t)dn * W + w_i]; }
    if (lt + rt + ut + dt <= 0.0f) return;
    const float l = lt > 0.0f ? 1.f : 0.f, r = rt > 0.0f ? 1.f : 0.f;
    const float u = ut > 0.0f ? 1.f : 0.f, d = dt > 0.0f ? 1.f : 0.f;
    const float den = l + r + u + d;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        float *o = out + ((size_t)b * 2 + ch) * HW;
        // volatile-free plain loads: the sources are non-hole pixels, final since the averaging pass
        const float v = l * o[(size_t)h_i * W + lo] + r * o[(size_t)h_i * W + ro] +
                        u * o[(size_t)uo * W + w_i] + d * o[(size_t)dn * W + w_i];
        o[(size_t)h_i * W + w_i] = v / den;
    }
}

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
        const size_t a[4] = {(size_t)c.T * W + c.L, (size_t)c.T * W + c.R, (size_t)c.Bm * W + c.L, (size_t)c.Bm * W + c.R};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float cnt = __ldg(cn + a[k]);
            const float gU = __ldg(gu + a[k]), gV = __ldg(gv + a[k]);
            if (DEPTH) {
                su += -gU * d / cnt;                                   // depth :289-296
                sv += -gV * d / cnt;
                sd += -gU / cnt * (fx - __ldg(ou + a[k]));             // depth :311-322
                sd += -gV / cnt * (fy - __ldg(ou + HW + a[k]));        // depth :324-335
            } else {
                su += -gU / cnt;                                       // :277-284
                sv += -gV / cnt;
            }
        }
    }
    st_stream(gi1 + ((size_t)b * 2 + 0) * HW + pix, su);
    st_stream(gi1 + ((size_t)b * 2 + 1) * HW + pix, sv);
    if (DEPTH) st_stream(gi2 + (size_t)b * HW + pix, sd);
}

template <bool DEPTH>
int projection_forward(const float *flow, const float *depth, float *count, float *out,
                       int B, int H, int W, int fillhole, cudaStream_t s)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !flow || !count || !out || (DEPTH && !depth)) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    const size_t HW = (size_t)H * W;
    int e = set_error(cudaMemsetAsync(count, 0, sizeof(float) * B * HW, s), "clear count");
    if (e) return e;
    e = set_error(cudaMemsetAsync(out, 0, sizeof(float) * 2 * B * HW, s), "clear output");
    if (e) return e;
    dim3 block(BX, BY), grid(ceil_div(W, BX), ceil_div(H, BY), B);
    projection_splat_kernel<DEPTH><<<grid, block, 0, s>>>(flow, depth, count, out, H, W);
    const long long rows_total = (long long)B * H;
    if (rows_total >= (1ll << 31)) return VFIDKR_ERR_ARG;
    const unsigned nb = (unsigned)std::min<long long>(rows_total, (long long)sm_count() * 8);
    projection_finish_kernel<<<nb, ROW_THREADS, 0, s>>>(count, out, H, W, (int)rows_total);
    note_launch(2);
    if (fillhole) {
        projection_fillhole_kernel<<<grid, block, 0, s>>>(count, out, H, W);
        note_launch();
    }
    return check_launch("flow projection forward");
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
{ return projection_forward<false>(input1, nullptr, count, output, B, H, W, fillhole, (cudaStream_t)s); }

VFIDKR_API int vfidkr_flowprojection_backward(const float *input1, const float *count, const float *gradoutput,
                                              float *gradinput1, int B, int H, int W, vfidkr_stream_t s)
{ return projection_backward<false>(input1, nullptr, count, nullptr, gradoutput, gradinput1, nullptr, B, H, W, (cudaStream_t)s); }

VFIDKR_API int vfidkr_depthflowprojection_forward(const float *input1, const float *input2, float *count,
                                                  float *output, int B, int H, int W, int fillhole,
                                                  vfidkr_stream_t s)
{ return projection_forward<true>(input1, input2, count, output, B, H, W, fillhole, (cudaStream_t)s); }

VFIDKR_API int vfidkr_depthflowprojection_backward(const float *input1, const float *input2, const float *count,
                                                   const float *output, const float *gradoutput,
                                                   float *gradinput1, float *gradinput2,
                                                   int B, int H, int W, vfidkr_stream_t s)
{ return projection_backward<true>(input1, input2, count, output, gradoutput, gradinput1, gradinput2, B, H, W, (cudaStream_t)s); }
