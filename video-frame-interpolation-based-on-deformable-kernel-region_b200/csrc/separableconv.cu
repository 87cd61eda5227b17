// separableconv.cu -- SeparableConv (per-pixel separable F x F filtering on the valid region) and
// SeparableConvFlow (filter centroid -> flow), forward and backward, for sm_100a.
//
// Behaviour follows my_package/SeparableConv/separableconv_cuda_kernel.cu:29-135 and
// my_package/SeparableConvFlow/separableconvflow_cuda_kernel.cu:29-174.
// Differences in HOW:
//   * SeparableConv backward uses no atomics at all.  The reference issues 3*C*F*F atomicAdds per
//     pixel (:122-127); here gradinput2/gradinput3 are thread-private register sums, and gradinput1
//     is computed as a GATHER over the (at most F x F) output pixels whose window covers the input
//     pixel -- deterministic and race-free.
//   * every output element is written, so no caller zero-fill is needed.
#include "common.cuh"

namespace vfidkr {
namespace {

constexpr int BX = 32, BY = 8;

// out[b,c,h,w] = sum_y sum_x I[b,c,h+y,w+x] * v[b,y,h,w] * hz[b,x,h,w]      (:65-77)
__global__ void __launch_bounds__(BX *BY)
sepconv_forward_kernel(const float *__restrict__ in1, const float *__restrict__ in2, const float *__restrict__ in3,
                       float *__restrict__ out, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= Wo || h_i >= Ho) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo, po = (size_t)h_i * Wo + w_i;
    const float *v = in2 + (size_t)b * F * HWo + po, *hz = in3 + (size_t)b * F * HWo + po;
    for (int c0 = 0; c0 < C; c0 += 3) {   // three channels per pass share the v/hz loads
        float acc[3] = {0.f, 0.f, 0.f};
        const float *pl = in1 + ((size_t)b * C + c0) * HW + (size_t)h_i * W + w_i;
        for (int y = 0; y < F; ++y) {
            const float vy = __ldg(v + (size_t)y * HWo);
            const float *row = pl + (size_t)y * W;
            for (int x = 0; x < F; ++x) {
                const float hx = __ldg(hz + (size_t)x * HWo);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    if (c0 + cc < C) acc[cc] += __ldg(row + (size_t)cc * HW + x) * vy * hx;   // temp1*temp2*temp3 (:76)
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            if (c0 + cc < C) out[((size_t)b * C + c0 + cc) * HWo + po] = acc[cc];
    }
}

// gradinput2[y] = sum_c g_c sum_x I*hz[x]   (:124-125);  gradinput3[x] = sum_c g_c sum_y I*v[y]   (:126-127)
__global__ void __launch_bounds__(BX *BY)
sepconv_backward_filters_kernel(const float *__restrict__ in1, const float *__restrict__ in2,
                                const float *__restrict__ in3, const float *__restrict__ gout,
                                float *__restrict__ gi2, float *__restrict__ gi3, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int w_i = blockIdx.x * BX + threadIdx.x, h_i = blockIdx.y * BY + threadIdx.y;
    if (w_i >= Wo || h_i >= Ho) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo, po = (size_t)h_i * Wo + w_i;
    const float *v = in2 + (size_t)b * F * HWo + po, *hz = in3 + (size_t)b * F * HWo + po;
    const float *g = gout + (size_t)b * C * HWo + po;
    const float *pl = in1 + (size_t)b * C * HW + (size_t)h_i * W + w_i;
    for (int y = 0; y < F; ++y) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gc = __ldg(g + (size_t)c * HWo);
            const float *row = pl + (size_t)c * HW + (size_t)y * W;
            float t = 0.f;
            for (int x = 0; x < F; ++x) t += __ldg(row + x) * __ldg(hz + (size_t)x * HWo);
            s += gc * t;
        }
        gi2[((size_t)b * F + y) * HWo + po] = s;
    }
    for (int x = 0; x < F; ++x) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gc = __ldg(g + (size_t)c * HWo);
            const float *col = pl + (size_t)c * HW + x;
            float t = 0.f;
            for (int y = 0; y < F; ++y) t += __ldg(col + (size_t)y * W) * __ldg(v + (size_t)y * HWo);
            s += gc * t;
        }
        gi3[((size_t)b * F + x) * HWo + po] = s;
    }
}

// gradinput1[c,Y,X] = sum over (y,x) with (Y-y, X-x) a valid output pixel of g[c]*v[y]*hz[x] there (:122-123)
__global__ void __launch_bounds__(BX *BY)
sepconv_backward_image_kernel(const float *__restrict__ in2, const float *__restrict__ in3,
                              const float *__restrict__ gout, float *__restrict__ gi1, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    const int X = blockIdx.x * BX + threadIdx.x, Y = blockIdx.y * BY + threadIdx.y;
    if (X >= W || Y >= H) return;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    const int y_lo = max(0, Y - (Ho - 1)), y_hi = min(F - 1, Y);
    const int x_lo = max(0, X - (Wo - 1)), x_hi = min(F - 1, X);
    for (int c0 = 0; c0 < C; c0 += 3) {
        float acc[3] = {0.f, 0.f, 0.f};
        for (int y = y_lo; y <= y_hi; ++y) {
            const size_t rowo = (size_t)(Y - y) * Wo;
            for (int x = x_lo; x <= x_hi; ++x) {
                const size_t po = rowo + (X - x);
                const float k = __ldg(in2 + ((size_t)b * F + y) * HWo + po) * __ldg(in3 + ((size_t)b * F + x) * HWo + po);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    if (c0 + cc < C) acc[cc] += __ldg(gout + ((size_t)b * C + c0 + cc) * HWo + po) * k;
            }
        }
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            if (c0 + cc < C) gi1[((size_t)b * C + c0 + cc) * HW + (size_t)Y * W + X] = acc[cc];
    }
}

// SeparableConvFlow: flow = centroid of the weights - (F-1)/2, sentinel -2000 when |sum| == 0 (:56-89)
__global__ void __launch_bounds__(256)
sepconvflow_forward_kernel(const float *__restrict__ in2, const float *__restrict__ in3, float *__restrict__ flow,
                           size_t HWo, size_t total, int F)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / HWo, po = idx - b * HWo;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {   // channel 0 <- input3 (x), channel 1 <- input2 (y)
            const float *src = (ch == 0 ? in3 : in2) + b * F * HWo + po;
            float num = 0.f, den = 0.f;
            for (int k = 0; k < F; ++k) {
                const float t = ld_stream(src + (size_t)k * HWo);
                num += (float)k * t;   // :61
                den += t;              // :62
            }
            // flow_y / sum_weights - ((float)(filter_size)-1.0)/2.0 is evaluated in double (:66)
            const float val = (float)((double)(num / den) - ((double)(float)F - 1.0) / 2.0);
            st_stream(flow + (b * 2 + ch) * HWo + po, fabsf(den) > 0.0f ? val : -2000.0f);
        }
    }
}

// gi[k] = g * (k / S - num / S^2) where |S| > 0, else 0 (:131-169)
__global__ void __launch_bounds__(256)
sepconvflow_backward_kernel(const float *__restrict__ in2, const float *__restrict__ in3,
                            const float *__restrict__ gflow, float *__restrict__ gi2, float *__restrict__ gi3,
                            size_t HWo, size_t total, int F)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / HWo, po = idx - b * HWo;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const float *src = (ch == 0 ? in3 : in2) + b * F * HWo + po;
            float *dst = (ch == 0 ? gi3 : gi2) + b * F * HWo + po;
            float num = 0.f, den = 0.f;
            for (int k = 0; k < F; ++k) {
                const float t = __ldg(src + (size_t)k * HWo);
                num += (float)k * t;
                den += t;
            }
            if (fabsf(den) > 0.0f) {
                const float g = ld_stream(gflow + (b * 2 + ch) * HWo + po);
                const float offset = num / (den * den);   // :138
                for (int k = 0; k < F; ++k) st_stream(dst + (size_t)k * HWo, g * ((float)k / den - offset));   // :140-143
            } else {
                for (int k = 0; k < F; ++k) st_stream(dst + (size_t)k * HWo, 0.0f);
            }
        }
    }
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

VFIDKR_API int vfidkr_separableconv_forward(const float *input1, const float *input2, const float *input3,
                                            float *output, int B, int C, int H, int W, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || F <= 0 || H - F + 1 <= 0 || W - F + 1 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !input3 || !output) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    dim3 block(BX, BY), grid(ceil_div(W - F + 1, BX), ceil_div(H - F + 1, BY), B);
    sepconv_forward_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(input1, input2, input3, output, C, H, W, F);
    note_launch();
    return check_launch("separableconv forward");
}

VFIDKR_API int vfidkr_separableconv_backward(const float *input1, const float *input2, const float *input3,
                                             const float *gradoutput, float *gradinput1, float *gradinput2,
                                             float *gradinput3, int B, int C, int H, int W, int F,
                                             vfidkr_stream_t stream)
{
    if (B <= 0 || C <= 0 || F <= 0 || H - F + 1 <= 0 || W - F + 1 <= 0 || B > 65535) return VFIDKR_ERR_ARG;
    if (!input1 || !input2 || !input3 || !gradoutput || !gradinput1 || !gradinput2 || !gradinput3) return VFIDKR_ERR_ARG;
    if ((long long)H * W >= (1ll << 31)) return VFIDKR_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(BX, BY), grid_o(ceil_div(W - F + 1, BX), ceil_div(H - F + 1, BY), B), grid_i(ceil_div(W, BX), ceil_div(H, BY), B);
    sepconv_backward_filters_kernel<<<grid_o, block, 0, s>>>(input1, input2, input3, gradoutput, gradinput2, gradinput3, C, H, W, F);
    sepconv_backward_image_kernel<<<grid_i, block, 0, s>>>(input2, input3, gradoutput, gradinput1, C, H, W, F);
    note_launch(2);
    return check_launch("separableconv backward");
}

VFIDKR_API int vfidkr_separableconvflow_forward(const float *input2, const float *input3, float *flow_output,
                                                int B, int Ho, int Wo, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || Ho <= 0 || Wo <= 0 || F <= 0 || !input2 || !input3 || !flow_output) return VFIDKR_ERR_ARG;
    const size_t HWo = (size_t)Ho * Wo, total = (size_t)B * HWo;
    const unsigned nb = (unsigned)min((size_t)sm_count() * 8, (total + 255) / 256);
    sepconvflow_forward_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(input2, input3, flow_output, HWo, total, F);
    note_launch();
    return check_launch("separableconvflow forward");
}

VFIDKR_API int vfidkr_separableconvflow_backward(const float *input2, const float *input3,
                                                 const float *gradflow_output, float *gradinput2, float *gradinput3,
                                                 int B, int Ho, int Wo, int F, vfidkr_stream_t stream)
{
    if (B <= 0 || Ho <= 0 || Wo <= 0 || F <= 0) return VFIDKR_ERR_ARG;
    if (!input2 || !input3 || !gradflow_output || !gradinput2 || !gradinput3) return VFIDKR_ERR_ARG;
    const size_t HWo = (size_t)Ho * Wo, total = (size_t)B * HWo;
    const unsigned nb = (unsigned)min((size_t)sm_count() * 8, (total + 255) / 256);
    sepconvflow_backward_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(input2, input3, gradflow_output, gradinput2, gradinput3, HWo, total, F);
    note_launch();
    return check_launch("separableconvflow backward");
}
