// frameio.cu -- frame I/O boundary of the demo drivers: uint8 HWC frames <-> padded float NCHW.
//
// Behaviour follows demo_MiddleBury.py:276-364 (colab_interpolate.py:85-148 is the same):
//   in : float32(u8) / 255.0, HWC -> CHW (:276-277), ReplicationPad2d([left, right, top, bottom]) (:303-309) where each
//        dimension is padded to the next multiple of 128, or by 32 + 32 if it already is one (:286-301);
//   out: crop [top : top + H, left : left + W] (:350-351), 255.0 * clip(y, 0, 1), round half to even (np.round), uint8 (:364).
// Byte / index work: bit-exact against numpy.  One thread per output element; reads are clamped gathers (pad) or a
// strided crop, writes are coalesced (channel-innermost for the uint8 frame).
#include "common.cuh"

namespace vfidkr {
namespace {

__global__ void __launch_bounds__(256)
frames_pad_kernel(const unsigned char *__restrict__ in, float *__restrict__ out, int H, int W, int Hp, int Wp, int top, int left)
{
    const int xp = blockIdx.x * 32 + threadIdx.x, yp = blockIdx.y * 8 + threadIdx.y;
    if (xp >= Wp || yp >= Hp) return;
    const int b = blockIdx.z;
    const int y = clampi(yp - top, 0, H - 1), x = clampi(xp - left, 0, W - 1);   // replication padding
    const unsigned char *px = in + (((size_t)b * H + y) * W + x) * 3;
    const size_t plane = (size_t)Hp * Wp, o = (size_t)b * 3 * plane + (size_t)yp * Wp + xp;
#pragma unroll
    for (int c = 0; c < 3; ++c) st_stream(out + o + (size_t)c * plane, __fdiv_rn((float)px[c], 255.0f));   // :276
}

__global__ void __launch_bounds__(256)
frames_crop_kernel(const float *__restrict__ in, unsigned char *__restrict__ out, int H, int W, int Hp, int Wp, int top, int left)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int b = blockIdx.z;
    const size_t plane = (size_t)Hp * Wp, i = (size_t)b * 3 * plane + (size_t)(y + top) * Wp + (x + left);
    unsigned char *px = out + (((size_t)b * H + y) * W + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = fminf(fmaxf(__ldcs(in + i + (size_t)c * plane), 0.0f), 1.0f);   // clip(0, 1) (:351); NaN -> 0
        px[c] = (unsigned char)rintf(__fmul_rn(255.0f, v));                            // np.round: half to even (:364)
    }
}

}  // namespace
}  // namespace vfidkr

using namespace vfidkr;

// padding of one dimension (demo_MiddleBury.py:286-301): returns the leading pad, *padded = the padded size
VFIDKR_API int vfidkr_frame_padding(int size, int *padded)
{
    int lead, total;
    if (size != ((size >> 7) << 7)) {
        total = (((size >> 7) + 1) << 7) - size;
        lead = total / 2;
    } else {
        total = 64;
        lead = 32;
    }
    if (padded) *padded = size + total;
    return lead;
}

VFIDKR_API int vfidkr_frames_u8_to_padded_f32(const unsigned char *frames, float *output, int B, int H, int W,
                                              vfidkr_stream_t stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !frames || !output) return VFIDKR_ERR_ARG;
    int Hp, Wp;
    const int top = vfidkr_frame_padding(H, &Hp), left = vfidkr_frame_padding(W, &Wp);
    if (ceil_div(Hp, 8) > 65535u) return VFIDKR_ERR_ARG;
    dim3 block(32, 8), grid(ceil_div(Wp, 32), ceil_div(Hp, 8), B);
    frames_pad_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(frames, output, H, W, Hp, Wp, top, left);
    note_launch();
    return check_launch("frames -> padded float");
}

VFIDKR_API int vfidkr_padded_f32_to_frames_u8(const float *padded, unsigned char *frames, int B, int H, int W,
                                              vfidkr_stream_t stream)
{
    if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || !frames || !padded) return VFIDKR_ERR_ARG;
    int Hp, Wp;
    const int top = vfidkr_frame_padding(H, &Hp), left = vfidkr_frame_padding(W, &Wp);
    if (ceil_div(H, 8) > 65535u) return VFIDKR_ERR_ARG;
    dim3 block(32, 8), grid(ceil_div(W, 32), ceil_div(H, 8), B);
    frames_crop_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(padded, frames, H, W, Hp, Wp, top, left);
    note_launch();
    return check_launch("padded float -> frames");
}
