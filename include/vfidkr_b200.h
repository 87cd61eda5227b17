/*
 * vfidkr_b200.h -- C ABI of libvfidkr_b200.so: hand-written sm_100a CUDA kernels for the
 * per-pixel sampling / warping hot path of VFIDKR.
 *
 * This is the drop-in boundary.  Each entry point replaces one symbol of the reference's
 * pybind11 extension modules (all citations relative to /root/reference/).  The reference
 * symbols take at::Tensor&; these take the same tensors as raw DEVICE pointers + sizes, so
 * the library has no torch (or any other framework) type in its signatures.
 *
 * Common contract
 *   - All tensors are float32, NCHW, contiguous (the reference asserts is_contiguous() in
 *     Python, e.g. FilterInterpolationLayer.py:16-18, and checks W-stride == 1 in C++,
 *     filterinterpolation_cuda.cc:579-581).
 *   - float32 ONLY.  The reference dispatches float and double (AT_DISPATCH_FLOATING_TYPES in
 *     every launcher, e.g. filterinterpolation_cuda_kernel.cu:3155) and half in the correlation
 *     (correlation_cuda_kernel.cu:386), although every intermediate in its kernels is `float`;
 *     VFIDKR itself only ever passes float32.  The Python front-end raises TypeError on any
 *     other dtype instead of converting silently.
 *   - The caller owns every INPUT and OUTPUT buffer and the library never synchronises; work is
 *     enqueued on `stream` (the reference uses at::cuda::getCurrentCUDAStream(),
 *     filterinterpolation_cuda.cc:590).  Some entry points need device SCRATCH memory (the
 *     projections' splat image, the correlation's split-K partial volumes and row-padded
 *     copies, the strip kernels' work queue): it is taken from a library-private, stream-ordered
 *     memory pool (cudaMallocFromPoolAsync on `stream`, freed on `stream` before the call
 *     returns) -- the counterpart of the rbot1/rbot2 scratch tensors the reference's
 *     correlation resizes inside C++ (correlation_cuda.cc:34-40).  The pool keeps at most
 *     VFIDKR_SCRATCH_RETAIN_MB (default 1024) MiB cached between calls; vfidkr_trim_scratch()
 *     returns everything to the device.  The process-wide default pool is not touched.
 *   - Unlike the reference, outputs need NOT be zero-filled by the caller: every element
 *     of every output/gradient buffer is written (accumulation targets are cleared on the
 *     stream by the library itself).
 *   - Return value: 0 on success; VFIDKR_ERR_ARG (1) for an argument the reference's .cc
 *     glue rejects with `return error` (shape mismatch); VFIDKR_ERR_CUDA (2) when the launch
 *     failed (the reference raises AT_ERROR("CUDA call failed") there).  The reference's
 *     correlation module inverts the convention (1 = success, correlation_cuda_kernel.cu:417-426);
 *     this ABI does not.
 *   - Thread-safe: no global mutable state except a relaxed launch counter, the per-device scratch
 *     pool and the process-wide test hooks below.
 *   - Environment (each read once, at first use): VFIDKR_SCRATCH_RETAIN_MB (above);
 *     VFIDKR_FI_FWD_PATH = strip | tile | direct forces a FilterInterpolation forward
 *     implementation (same as vfidkr_debug_force_forward_path); VFIDKR_FI_STRIP_TW = 128 | 144
 *     forces the tile width of the "_ori" strip kernel (experiments; the default takes 144 columns
 *     wherever the image is at least 192 wide).  Debug builds: -DVFIDKR_STRIP_STATS (pipeline
 *     counters), -DVFIDKR_BOUNDS_CHECK (every shared-memory window tap verified against global
 *     memory; vfidkr_debug_bounds_counts) -- neither symbol exists in a production build.
 */
#ifndef VFIDKR_B200_H_
#define VFIDKR_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without pulling in the CUDA headers */
typedef struct CUstream_st *vfidkr_stream_t;

#define VFIDKR_OK 0
#define VFIDKR_ERR_ARG 1
#define VFIDKR_ERR_CUDA 2

/* ---- library info ------------------------------------------------------------------- */
/* ABI version (major*100+minor). */
int vfidkr_abi_version(void);
/* Number of kernel launches issued by this library since load (all threads). */
unsigned long long vfidkr_launch_count(void);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char *vfidkr_last_error(void);
/* TEST HOOK (no reference counterpart): force the implementation of the FilterInterpolation forwards --
 * 0 = automatic (production), 1 = strip kernels, 2 = tile kernel, 3 = direct kernels -- so that the parity tests
 * can hold every implementation to the oracle.  Process-wide; returns the previous setting, -1 on a bad value. */
int vfidkr_debug_force_forward_path(int path);
/* TEST / MEASUREMENT HOOK: implementation of the correlation forward (kernel_size 1, strides 1, max_displacement 4, pad 4) --
 * 0 = automatic, 1 = the SIMT register-tile kernels, 2 = the tensor-core kernel (tcgen05.mma kind::tf32, 3 x TF32 split,
 * accumulator in tensor memory; correlation_tc.cu).  Process-wide; returns the previous setting, -1 on a bad value. */
int vfidkr_debug_force_correlation_path(int path);
/* Return the scratch memory the library's private pool has cached on the current device to the device
 * (see "Common contract"; blocks still in use by enqueued work are not affected).  No reference counterpart. */
int vfidkr_trim_scratch(void);

/* ---- FilterInterpolation: adaptive warping with per-pixel F x F filters ------------------
 * input1 [B,C,H,W] image/features   input2 [B,2,H,W] flow (x,y)   input3 [B,F*F,H,W] filter
 * input4 [B,2*F*F,H,W] DKR offset field (y-offsets in channels [0,F*F), x-offsets after)
 * filter_size F: any F >= 1 (fast path for F == 4, the only size VFIDKR uses).
 */

/* replaces filterinterpolation_cuda.FilterInterpolationLayer_gpu_forward_ori
 * (filterinterpolation_cuda.cc:537-606; kernel filterinterpolation_cuda_kernel.cu:2692-2823) */
int vfidkr_filterinterpolation_forward_ori(const float *input1, const float *input2, const float *input3,
                                           float *output, int B, int C, int H, int W, int filter_size,
                                           vfidkr_stream_t stream);
/* The same forward with a blend epilogue: output = scale * FI(input1, input2, input3) (+ output's previous contents when
 * accumulate != 0).  Two calls warp both directions and blend them -- ref0/2 + ref2/2 (networks/DAIN.py:573),
 * (1-t)*ref0 + t*ref2 (DAIN_slowmotion.py:335) -- without the two intermediate frames and the blend pass.
 * out_batch_stride (elements; 0 = dense, C*H*W) lets `output` be a channel slice of a wider tensor, e.g. the
 * rectify-input concat of networks/DAIN.py:264-269: channels stay H*W apart, batch items out_batch_stride apart.
 * The reference has no such symbol; scale = 1, accumulate = 0, out_batch_stride = 0 is
 * vfidkr_filterinterpolation_forward_ori. */
int vfidkr_filterinterpolation_forward_ori_blend(const float *input1, const float *input2, const float *input3,
                                                 float *output, int B, int C, int H, int W, int filter_size,
                                                 float scale, int accumulate, long long out_batch_stride,
                                                 vfidkr_stream_t stream);
/* replaces ..._gpu_backward_ori (filterinterpolation_cuda.cc:608-687; kernel :2827-3125) */
int vfidkr_filterinterpolation_backward_ori(const float *input1, const float *input2, const float *input3,
                                            const float *gradoutput, float *gradinput1, float *gradinput2,
                                            float *gradinput3, int B, int C, int H, int W, int filter_size,
                                            vfidkr_stream_t stream);

/* replaces FilterInterpolationLayer_gpu_forward / _gpu_backward, the 4-input DKR family with
 * static quadrants (filterinterpolation_cuda.cc:11-187; kernels :29-426, :430-1215).
 * As in the reference the forward only computes for F in {4,6} (:68) and writes zeros otherwise. */
int vfidkr_filterinterpolation_forward_dkr(const float *input1, const float *input2, const float *input3,
                                           const float *input4, float *output,
                                           int B, int C, int H, int W, int filter_size, vfidkr_stream_t stream);
int vfidkr_filterinterpolation_backward_dkr(const float *input1, const float *input2, const float *input3,
                                            const float *input4, const float *gradoutput,
                                            float *gradinput1, float *gradinput2, float *gradinput3,
                                            float *gradinput4, int B, int C, int H, int W, int filter_size,
                                            vfidkr_stream_t stream);

/* replaces ..._gpu_forward_deforconv / _gpu_backward_deforconv, DKR with data-dependent quadrants
 * (filterinterpolation_cuda.cc:191-367; kernels :1353-1496, :1500-1935) */
int vfidkr_filterinterpolation_forward_deforconv(const float *input1, const float *input2, const float *input3,
                                                 const float *input4, float *output,
                                                 int B, int C, int H, int W, int filter_size,
                                                 vfidkr_stream_t stream);
int vfidkr_filterinterpolation_backward_deforconv(const float *input1, const float *input2, const float *input3,
                                                  const float *input4, const float *gradoutput,
                                                  float *gradinput1, float *gradinput2, float *gradinput3,
                                                  float *gradinput4, int B, int C, int H, int W, int filter_size,
                                                  vfidkr_stream_t stream);

/* replaces ..._gpu_forward_nofilterwithdeforconv / _gpu_backward_nofilterwithdeforconv
 * (filterinterpolation_cuda.cc:374-533; kernels :2070-2191, :2195-2567).
 * Here input3 / gradinput3 are the [B,2*F*F,H,W] offset field and its gradient. */
int vfidkr_filterinterpolation_forward_nofilterwithdeforconv(const float *input1, const float *input2,
                                                             const float *input3, float *output,
                                                             int B, int C, int H, int W, int filter_size,
                                                             vfidkr_stream_t stream);
int vfidkr_filterinterpolation_backward_nofilterwithdeforconv(const float *input1, const float *input2,
                                                              const float *input3, const float *gradoutput,
                                                              float *gradinput1, float *gradinput2,
                                                              float *gradinput3, int B, int C, int H, int W,
                                                              int filter_size, vfidkr_stream_t stream);

/* ---- FlowProjection / DepthFlowProjection: forward splat of -flow with atomics -----------
 * input1 [B,2,H,W] flow; input2 [B,1,H,W] inverse depth (> 0); count [B,1,H,W]; output [B,2,H,W].
 * fillhole != 0 runs the hole-filling pass (inference; FlowProjectionLayer.py:23).
 */
/* replaces flowprojection_cuda.FlowProjectionLayer_gpu_forward / _gpu_backward
 * (flowprojection_cuda.cc:9-114; kernels flowprojection_cuda_kernel.cu:29-301) */
int vfidkr_flowprojection_forward(const float *input1, float *count, float *output,
                                  int B, int H, int W, int fillhole, vfidkr_stream_t stream);
int vfidkr_flowprojection_backward(const float *input1, const float *count, const float *gradoutput,
                                   float *gradinput1, int B, int H, int W, vfidkr_stream_t stream);
/* replaces depthflowprojection_cuda.DepthFlowProjectionLayer_gpu_forward / _gpu_backward
 * (depthflowprojection_cuda.cc:10-143; kernels depthflowprojection_cuda_kernel.cu:29-341) */
int vfidkr_depthflowprojection_forward(const float *input1, const float *input2, float *count, float *output,
                                       int B, int H, int W, int fillhole, vfidkr_stream_t stream);
int vfidkr_depthflowprojection_backward(const float *input1, const float *input2, const float *count,
                                        const float *output, const float *gradoutput,
                                        float *gradinput1, float *gradinput2,
                                        int B, int H, int W, vfidkr_stream_t stream);

/* ---- Interpolation / InterpolationCh: bilinear backward warp, zero fill ---------------------
 * replaces interpolation_cuda.InterpolationLayer_gpu_forward / _gpu_backward
 * (interpolation_cuda.cc:10-121; kernels interpolation_cuda_kernel.cu:29-204) and the identical
 * interpolationch_cuda twins.  `require_c3` != 0 reproduces Interpolation's C == 3 check
 * (interpolation_cuda.cc:19); InterpolationCh passes 0.
 */
int vfidkr_interpolation_forward(const float *input1, const float *input2, float *output,
                                 int B, int C, int H, int W, int require_c3, vfidkr_stream_t stream);
int vfidkr_interpolation_backward(const float *input1, const float *input2, const float *gradoutput,
                                  float *gradinput1, float *gradinput2,
                                  int B, int C, int H, int W, int require_c3, vfidkr_stream_t stream);

/* ---- SeparableConv / SeparableConvFlow ---------------------------------------------------------
 * input1 [B,C,H,W]; input2 (vertical) / input3 (horizontal) [B,F,H-F+1,W-F+1].
 * replaces separableconv_cuda.SeparableConvLayer_gpu_forward / _gpu_backward
 * (separableconv_cuda.cc:10-176, C == 3 enforced at :21; kernels separableconv_cuda_kernel.cu:29-135)
 * backward: every element of the three gradients is written; gradinput1 is cleared on the stream by the library
 * and summed with atomics (about 9 per element for filter sizes that use the shared-memory tiled kernels), so its
 * low bits depend on the order of arrival, as the reference's do (F*F atomics per element there).
 */
int vfidkr_separableconv_forward(const float *input1, const float *input2, const float *input3, float *output,
                                 int B, int C, int H, int W, int filter_size, vfidkr_stream_t stream);
int vfidkr_separableconv_backward(const float *input1, const float *input2, const float *input3,
                                  const float *gradoutput, float *gradinput1, float *gradinput2,
                                  float *gradinput3, int B, int C, int H, int W, int filter_size,
                                  vfidkr_stream_t stream);
/* replaces separableconvflow_cuda.SeparableConvFlowLayer_gpu_forward / _gpu_backward
 * (separableconvflow_cuda.cc; kernels separableconvflow_cuda_kernel.cu:29-174).
 * flow_output [B,2,Ho,Wo]; input1 only contributes its shape (H = Ho+F-1, W = Wo+F-1) and a zero gradient. */
int vfidkr_separableconvflow_forward(const float *input2, const float *input3, float *flow_output,
                                     int B, int Ho, int Wo, int filter_size, vfidkr_stream_t stream);
int vfidkr_separableconvflow_backward(const float *input2, const float *input3, const float *gradflow_output,
                                      float *gradinput2, float *gradinput3,
                                      int B, int Ho, int Wo, int filter_size, vfidkr_stream_t stream);

/* ---- Correlation (FlowNet / PWC-Net cost volume) -----------------------------------------------
 * replaces correlation_cuda.forward / backward (correlation_cuda.cc:8-165; kernels
 * correlation_cuda_kernel.cu:47-334).  Reads NCHW directly: the reference's rbot1/rbot2 padded
 * NHWC scratch tensors are not needed.  output [B,OC,OH,OW] with the shape rules of
 * correlation_cuda.cc:23-36 (use vfidkr_correlation_outshape).  corr_type_multiply is accepted
 * and ignored exactly as the reference ignores it.  Backward is in-contract for stride1 == 1
 * and pad_size >= max_displacement (the reference reads out of bounds otherwise).
 */
int vfidkr_correlation_outshape(int H, int W, int pad_size, int kernel_size, int max_displacement,
                                int stride1, int stride2, int *out_channels, int *out_h, int *out_w);
int vfidkr_correlation_forward(const float *input1, const float *input2, float *output,
                               int B, int C, int H, int W, int pad_size, int kernel_size,
                               int max_displacement, int stride1, int stride2, int corr_type_multiply,
                               vfidkr_stream_t stream);
/* No reference counterpart (it calls the correlation once per direction): both temporal directions of one pyramid
 * level in ONE launch -- output12 = correlation(input1, input2), output21 = correlation(input2, input1): the results of
 * two vfidkr_correlation_forward calls (bit-identical, except that maps small enough for the channel split may be split
 * differently, i.e. summed in another order).  Twice the tiles per launch: the coarse pyramid levels fill the machine. */
int vfidkr_correlation_forward_pair(const float *input1, const float *input2, float *output12, float *output21,
                                    int B, int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                                    int stride1, int stride2, int corr_type_multiply, vfidkr_stream_t stream);
int vfidkr_correlation_backward(const float *input1, const float *input2, const float *gradoutput,
                                float *gradinput1, float *gradinput2,
                                int B, int C, int H, int W, int pad_size, int kernel_size,
                                int max_displacement, int stride1, int stride2, int corr_type_multiply,
                                vfidkr_stream_t stream);

/* ---- flow pre-processing fused into the projection's read (networks/DAIN.py:306-308):
 * FlowProjection / DepthFlowProjection of  Upsample(x4, bilinear, align_corners = False)(scale0 * flow_lowres * scale1)
 * without materialising the full-resolution flow.  flow_lowres [B,2,h,w]; input2 [B,1,4h,4w] or NULL (FlowProjection);
 * count [B,1,4h,4w], output [B,2,4h,4w].  vfidkr_flow_upsample4 writes the enlarged flow itself (same arithmetic). ---- */
int vfidkr_flowprojection_forward_lowres(const float *flow_lowres, float scale0, float scale1, const float *input2,
                                         float *count, float *output, int B, int h, int w, int fillhole,
                                         vfidkr_stream_t stream);
int vfidkr_flow_upsample4(const float *flow_lowres, float scale0, float scale1, float *output, int B, int h, int w,
                          vfidkr_stream_t stream);
/* Adjoint of vfidkr_flow_upsample4 (what autograd's backward of scale -> nn.Upsample computes in the reference,
 * networks/DAIN.py:306-308): grad_output [B,2,4h,4w] -> grad_flow_lowres [B,2,h,w].  With the projection backward on the
 * enlarged flow this trains through vfidkr_flowprojection_forward_lowres (vfidkr_b200.flow_project_lowres). */
int vfidkr_flow_upsample4_backward(const float *grad_output, float scale0, float scale1, float *grad_flow_lowres,
                                   int B, int h, int w, vfidkr_stream_t stream);

/* ---- MinDepthFlowProjection (my_package/MinDepthFlowProjection/mindepthflowprojection_cuda.cc; kernels
 * mindepthflowprojection_cuda_kernel.cu:29-312): the in-range source pixel with the largest input2 wins its top-left
 * cell, output = -flow of the winner, count = its input2; hole filling as in the other projections.  The reference's
 * forward is a non-atomic read-compare-write and therefore timing-dependent; this is the deterministic version of the
 * same rule (64-bit atomicMax; ties go to the lowest pixel index).  input1 [B,2,H,W], input2 [B,1,H,W];
 * gradinput2 is written as zeros (the reference leaves the caller's zeros). ---- */
int vfidkr_mindepthflowprojection_forward(const float *input1, const float *input2, float *count, float *output,
                                          int B, int H, int W, int fillhole, vfidkr_stream_t stream);
int vfidkr_mindepthflowprojection_backward(const float *input1, const float *input2, const float *count,
                                           const float *gradoutput, float *gradinput1, float *gradinput2,
                                           int B, int H, int W, vfidkr_stream_t stream);

/* ---- PWCDCNet.warp (PWCNet/PWCNet.py:159-199): grid + flow, normalised with 2 v / max(size - 1, 1) - 1 (:178-179),
 * grid_sample (bilinear, zero padding) and the validity mask (grid_sample of ones, thresholded at 0.9999),
 * output = sample * mask.  align_corners != 0: grid_sample as in the reference's pinned torch 1.0.1
 * (environment.yaml:88,104), ix = x + fx -- the exact warp; align_corners == 0: what the same source line computes on
 * torch >= 1.3 (ix = (x + fx) W / (W - 1) - 0.5).  Not a native symbol of the reference: it replaces two grid_sample
 * launches and five elementwise kernels of PyTorch code.  x [B,C,H,W], flow [B,2,H,W].
 * The backward returns d/dx (scattered, cleared by the library) and d/dflow; the mask carries no gradient. ---- */
int vfidkr_pwcwarp_forward(const float *x, const float *flow, float *output, int B, int C, int H, int W,
                           int align_corners, vfidkr_stream_t stream);
int vfidkr_pwcwarp_backward(const float *x, const float *flow, const float *gradoutput, float *gradx, float *gradflow,
                            int B, int C, int H, int W, int align_corners, vfidkr_stream_t stream);

/* ---- frame I/O boundary of the demo drivers (demo_MiddleBury.py:276-364; not native in the reference: numpy + torch
 * ReplicationPad2d there).  frames: uint8 [B,H,W,3] (HWC) on the device.  padded: float32 [B,3,Hp,Wp] with
 * Hp / Wp from vfidkr_frame_padding (next multiple of 128, or + 64 if already one; leading pad returned).
 * u8 -> f32: v / 255, replication padding.  f32 -> u8: crop, 255 * clip(v, 0, 1), round half to even.  Bit-exact. ---- */
int vfidkr_frame_padding(int size, int *padded);
int vfidkr_frames_u8_to_padded_f32(const unsigned char *frames, float *padded, int B, int H, int W,
                                   vfidkr_stream_t stream);
int vfidkr_padded_f32_to_frames_u8(const float *padded, unsigned char *frames, int B, int H, int W,
                                   vfidkr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VFIDKR_B200_H_ */
