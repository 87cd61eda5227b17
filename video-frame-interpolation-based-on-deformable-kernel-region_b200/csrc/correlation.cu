// correlation.cu -- FlowNet / PWC-Net cost volume, forward and backward, for sm_100a.
//
// Behaviour follows PWCNet/correlation_package_pytorch1_0/correlation_cuda_kernel.cu:47-334 and the
// shape rules of correlation_cuda.cc:23-36.  What is different:
//   * NCHW is read directly with zero padding applied on the fly -- the reference's channels_first
//     repack into padded NHWC scratch tensors (two extra passes + two fill_ kernels) is gone;
//   * fast path for the only configuration VFIDKR uses (kernel_size 1, stride1 = stride2 = 1,
//     PWCNet.py:72): shared-memory tiles of both feature maps, each thread owns 4 adjacent pixels x
//     9 horizontal displacements of one vertical displacement, so every f2 value fetched from shared
//     memory feeds up to 4 FMAs (the reference runs 81 serial warp reductions per output pixel);
//   * backward for the fast path keeps the 81 upstream gradients of a pixel in registers and reuses
//     them for every channel, one launch for the whole batch (the reference launches per batch item);
//   * a generic path restates the reference formulas for any (kernel_size, stride1, stride2).
// fp32 FMA throughput, not HBM, bounds this op on SIMT (162*C flop vs 4*(2C+81) bytes per pixel);
// tcgen05 is not used: the 9-wide band of the (T+8)-wide product wastes >= 89% of a GEMM tile and
// 1e-5 parity needs a 3xTF32 split, which costs more tensor time than the SIMT kernel.
#include "common.cuh"

namespace vfidkr {
namespace {

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
// fast forward: kernel_size 1, stride1 = stride2 = 1, displacement radius DR (9 x 9 for DR = 4).
// Block = TY x TX output pixels; threads = (TX/4 pixel quads) x (2*DR+1 vertical displacements) x TY.
// ---------------------------------------------------------------------------------------------
template <int DR>
struct FastCfg {
    static constexpr int D = 2 * DR + 1;
    static constexpr int TX = 32, TY = 8, CK = 8;
    static constexpr int F2W = TX + 2 * DR, F2H = TY + 2 * DR;
    static constexpr int F2P = (F2W + 3) / 4 * 4;   // row pitch in floats (16-byte rows for LDS.128)
    static constexpr int THREADS = (TX / 4) * D * TY;
    static constexpr int SMEM_F1 = CK * TY * TX, SMEM_F2 = CK * F2H * F2P;
};

template <int DR>
__global__ void __launch_bounds__(FastCfg<DR>::THREADS)
corr_forward_fast_kernel(const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                         int C, int H, int W, int shift, int oh, int ow)
{
    using K = FastCfg<DR>;
    constexpr int D = K::D;
    __shared__ __align__(16) float s1[K::SMEM_F1];
    __shared__ __align__(16) float s2[K::SMEM_F2];

    const int n = blockIdx.z;
    const int ox0 = blockIdx.x * K::TX, oy0 = blockIdx.y * K::TY;   // output tile origin
    const int ix0 = ox0 + shift, iy0 = oy0 + shift;                 // same pixel in image coordinates
    const size_t HW = (size_t)H * W;
    const float *f1 = in1 + (size_t)n * C * HW, *f2 = in2 + (size_t)n * C * HW;

    const int tid = threadIdx.x;
    const int g = tid % (K::TX / 4);             // pixel quad inside the row
    const int tj = (tid / (K::TX / 4)) % D;      // vertical displacement index 0..D-1  (tj - DR)
    const int ry = tid / ((K::TX / 4) * D);      // row inside the tile

    float acc[4][D];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int t = 0; t < D; ++t) acc[p][t] = 0.0f;

    for (int c0 = 0; c0 < C; c0 += K::CK) {
        // ---- stage CK channels of both maps (zero outside the image) ----
        for (int idx = tid; idx < K::SMEM_F1; idx += K::THREADS) {
            const int x = idx % K::TX, y = (idx / K::TX) % K::TY, c = idx / (K::TX * K::TY);
            const int gy = iy0 + y, gx = ix0 + x;
            float v = 0.0f;
            if (c0 + c < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(f1 + (size_t)(c0 + c) * HW + (size_t)gy * W + gx);
            s1[idx] = v;
        }
        for (int idx = tid; idx < K::CK * K::F2H * K::F2W; idx += K::THREADS) {
            const int x = idx % K::F2W, y = (idx / K::F2W) % K::F2H, c = idx / (K::F2W * K::F2H);
            const int gy = iy0 - DR + y, gx = ix0 - DR + x;
            float v = 0.0f;
            if (c0 + c < C && gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(f2 + (size_t)(c0 + c) * HW + (size_t)gy * W + gx);
            s2[(c * K::F2H + y) * K::F2P + x] = v;
        }
        __syncthreads();
        // ---- 4 pixels x D horizontal displacements per thread ----
#pragma unroll
        for (int c = 0; c < K::CK; ++c) {
            const float4 a = *reinterpret_cast<const float4 *>(&s1[(c * K::TY + ry) * K::TX + 4 * g]);
            const float *row = &s2[(c * K::F2H + ry + tj) * K::F2P + 4 * g];
            float bv[4 + 2 * DR + 2];   // 12 for DR = 4 (three LDS.128)
#pragma unroll
            for (int q = 0; q < (4 + 2 * DR + 3) / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(row + 4 * q);
                bv[4 * q + 0] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w;
            }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int t = 0; t < D; ++t) acc[p][t] = fmaf(av[p], bv[p + t], acc[p][t]);
        }
        __syncthreads();
    }
    // ---- write: channel (tj, ti), row oy0+ry, pixels ox0+4g .. +3 ----
    const int oy = oy0 + ry, ox = ox0 + 4 * g;
    if (oy >= oh) return;
    const float inv = 1.0f / (float)C;   // nelems = kernel_size^2 * C (:104); the division is applied as acc / nelems
    const float nel = (float)C;
    (void)inv;
    const size_t plane = (size_t)oh * ow;
    float *o = out + ((size_t)n * D * D + (size_t)tj * D) * plane + (size_t)oy * ow + ox;
    const bool vec = (ox + 3 < ow) && ((((size_t)oy * ow + ox) & 3) == 0) && ((plane & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int t = 0; t < D; ++t) {
        float *ot = o + (size_t)t * plane;
        if (vec) {
            st_stream4(ot, make_float4(acc[0][t] / nel, acc[1][t] / nel, acc[2][t] / nel, acc[3][t] / nel));
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (ox + p < ow) ot[p] = acc[p][t] / nel;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fast backward (kernel_size 1, strides 1): one thread per input pixel keeps the D*D upstream
// gradients in registers and loops channels.
//   gi1[n,c,y,x] = 1/C * sum_tc g[n,tc,y+s,x+s]       * f2[n,c,y+tj,x+ti]          (:151-241)
//   gi2[n,c,y,x] = 1/C * sum_tc g[n,tc,y+s-tj,x+s-ti] * f1[n,c,y-tj,x-ti]          (:244-334)
// with s = pad - md, zero outside the image / the output plane.
// ---------------------------------------------------------------------------------------------
template <int DR, int WHICH>
__global__ void __launch_bounds__(256)
corr_backward_fast_kernel(const float *__restrict__ other, const float *__restrict__ gout, float *__restrict__ gi,
                          int C, int H, int W, int s, int oh, int ow)
{
    constexpr int D = 2 * DR + 1;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int n = blockIdx.z;
    const size_t HW = (size_t)H * W, plane = (size_t)oh * ow;
    const float *g = gout + (size_t)n * D * D * plane;
    float gv[D * D];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = 0; b < D; ++b) {
            const int tj = a - DR, ti = b - DR;
            const int oy = WHICH == 1 ? y + s : y + s - tj, ox = WHICH == 1 ? x + s : x + s - ti;
            gv[a * D + b] = (oy >= 0 && oy < oh && ox >= 0 && ox < ow) ? __ldg(g + (size_t)(a * D + b) * plane + (size_t)oy * ow + ox) : 0.0f;
        }
    const float nel = (float)C;
    for (int c = 0; c < C; ++c) {
        const float *pl = other + ((size_t)n * C + c) * HW;
        float acc = 0.0f;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const int yy = WHICH == 1 ? y + (a - DR) : y - (a - DR);
            const bool yin = yy >= 0 && yy < H;
#pragma unroll
            for (int b = 0; b < D; ++b) {
                const int xx = WHICH == 1 ? x + (b - DR) : x - (b - DR);
                const float v = (yin && xx >= 0 && xx < W) ? __ldg(pl + (size_t)yy * W + xx) : 0.0f;
                acc = fmaf(gv[a * D + b], v, acc);
            }
        }
        st_stream(gi + ((size_t)n * C + c) * HW + (size_t)y * W + x, acc / nel);
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
    if (k == 1 && s1 == 1 && s2 == 1 && md == 4) {
        using K = FastCfg<4>;
        dim3 grid(ceil_div(cs.ow, K::TX), ceil_div(cs.oh, K::TY), B);
        corr_forward_fast_kernel<4><<<grid, K::THREADS, 0, s>>>(input1, input2, output, C, H, W, md - pad, cs.oh, cs.ow);
    } else {
        dim3 block(32, 8), grid(ceil_div(cs.ow, 32), ceil_div(cs.oh, 8), B);
        corr_forward_generic_kernel<<<grid, block, 0, s>>>(input1, input2, output, C, H, W, pad, k, md, s1, s2, cs);
    }
    note_launch();
    return check_launch("correlation forward");
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
        dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B);
        corr_backward_fast_kernel<4, 1><<<grid, block, 0, s>>>(input2, gradoutput, gradinput1, C, H, W, pad - md, cs.oh, cs.ow);
        corr_backward_fast_kernel<4, 2><<<grid, block, 0, s>>>(input1, gradoutput, gradinput2, C, H, W, pad - md, cs.oh, cs.ow);
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
