// fi_bigc.cu -- FilterInterpolation "_ori" forward for MANY channels (C > 4; the 196-channel context warp of
// DAIN_slowmotion.py:311-317), sm_100a.
//
// What is computed follows my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu:2692-2823.
// With many channels the image dominates the traffic (8*C of the 8*C + 72 bytes per pixel) and the per-pixel work is
// 16 gathers + 16 FMAs per channel.  Through L1 those gathers bound the generic kernels at ~18 % of the HBM roofline
// (4 sectors per warp-wide LDG, tag lookups); the rolling-window strip kernel (fi_strip.cu) holds at most 4 channels.
// Here a CTA owns a 64 x 8 pixel tile (two pixels per thread, four rows apart; 64 x 16 tiles with 512 threads halve the
// region traffic but were measured 4-15 % slower: one CTA per SM leaves the barrier bubbles uncovered):
//   * flow, the 16 filter taps and the window geometry of both pixels are read ONCE and stay in registers for all
//     channels (the reference re-reads the taps per channel, :2755);
//   * a RW x RH region of the image, placed around the tile displaced by its MEAN integer flow, is streamed channel
//     group by channel group (CG = 4) into a two-stage shared-memory ring by one TMA box load per group; pixels whose
//     clamped 4 x 4 window lies inside the region (all of them unless the flow differs from the tile mean by more
//     than ~10 pixels) gather with LDS -- conflict-free: lanes are neighbouring pixels and the pitch is 3 x 32 words;
//   * the other pixels run the same arithmetic with clamped global gathers, pixel by pixel: correctness never depends
//     on the flow; out-of-range pixels copy input1 (:2814-2819).
// Preconditions (launcher; otherwise "not applicable"): F == 4, W % 4 == 0, 16-byte aligned image base.
#include <climits>

#include "common.cuh"
#include "fi_common.cuh"
#include "tma.cuh"

namespace vfidkr {
namespace {

namespace bigc {
constexpr int TW = 64, TH = 8, NT = 256, PX = 2;       // thread (tx, ty) owns pixels (tx, ty) and (tx, ty + 4)
constexpr int RW = 96, RH = 32, CG = 4, STAGES = 2;    // staged region per channel, channels per stage
constexpr int PLANE = RW * RH;
constexpr uint32_t STAGE_BYTES = CG * PLANE * sizeof(float);
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 128;
}  // namespace bigc

template <bool BLEND>
__global__ void __launch_bounds__(bigc::NT, 2)
fi_forward_ori_bigc_kernel(const __grid_constant__ CUtensorMap map_img, const float *__restrict__ in1,
                           const float *__restrict__ in2, const float *__restrict__ in3, float *__restrict__ out,
                           int C, int H, int W, float scale, int accumulate, size_t out_bs)
{
    // epilogue as in the other "_ori" forwards (vfidkr_filterinterpolation_forward_ori_blend): output = scale * result
    // (+ what output held), batch items of the output out_bs elements apart -- the warped context features of
    // DAIN_slowmotion.py:167-181 land directly in their channel slice of the 437-channel rectify input
    using namespace bigc;
    // (a separate instantiation: carried by the plain kernel as run-time flags the epilogue cost it 10 %)
    auto put = [&](float *dst, float v) {
        if (BLEND) {
            v *= scale;
            if (accumulate) v += __ldcs(dst);
        }
        st_stream(dst, v);
    };
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_img = reinterpret_cast<float *>(smem_raw);                       // [STAGES][CG][RH][RW]
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_img + STAGES * CG * PLANE);
    __shared__ int s_sum[3];   // sum of (L - x), sum of (T - y), number of in-range pixels

    const int tid = threadIdx.x, tx = tid % TW, ty = tid / TW, lane = tid & 31;
    const int b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const int w_i = blockIdx.x * TW + tx;

    if (tid == 0) {
        prefetch_tensormap(&map_img);
        for (int s = 0; s < STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
        s_sum[0] = 0; s_sum[1] = 0; s_sum[2] = 0;
    }
    __syncthreads();

    // ---- per-pixel state, channel independent ----
    bool inside[PX], active[PX];
    float fxy[PX][2];                  // the flow; the window geometry is re-derived from it where needed (registers)
    float w[PX][16], q[PX][4];
    size_t pix[PX];
    int sdx = 0, sdy = 0, sn = 0;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const int h_i = blockIdx.y * TH + ty + (TH / PX) * p;
        inside[p] = w_i < W && h_i < H;
        pix[p] = inside[p] ? (size_t)h_i * W + w_i : 0;
        float fx = 0.0f, fy = 0.0f;
        if (inside[p]) {
            fx = ld_stream(in2 + ((size_t)b * 2 + 0) * HW + pix[p]);
            fy = ld_stream(in2 + ((size_t)b * 2 + 1) * HW + pix[p]);
        }
        fxy[p][0] = fx; fxy[p][1] = fy;
        const FiPix fp = fi_pixel(w_i, h_i, fx, fy, W, H, 4);
        active[p] = inside[p] && fp.in_range;
#pragma unroll
        for (int k = 0; k < 16; ++k) w[p][k] = active[p] ? ld_stream(in3 + ((size_t)b * 16 + k) * HW + pix[p]) : 0.0f;
        q[p][0] = (1 - fp.alpha) * (1 - fp.beta); q[p][1] = fp.alpha * (1 - fp.beta);
        q[p][2] = (1 - fp.alpha) * fp.beta;       q[p][3] = fp.alpha * fp.beta;
        if (active[p]) { sdx += fp.L - w_i; sdy += fp.T - h_i; ++sn; }
    }
    sdx = __reduce_add_sync(0xffffffffu, sdx); sdy = __reduce_add_sync(0xffffffffu, sdy); sn = __reduce_add_sync(0xffffffffu, sn);
    if (lane == 0 && sn > 0) { atomicAdd(&s_sum[0], sdx); atomicAdd(&s_sum[1], sdy); atomicAdd(&s_sum[2], sn); }
    __syncthreads();
    const int n_act = s_sum[2];
    const bool any = n_act > 0;
    // region origin: the tile's window block (TW + 3 x TH + 3) displaced by the mean flow, centred in the region;
    // x on a 16-byte boundary (TMA box origins are kept at multiples of 4 elements)
    const int mdx = any ? (int)floorf((float)s_sum[0] / (float)n_act + 0.5f) : 0;
    const int mdy = any ? (int)floorf((float)s_sum[1] / (float)n_act + 0.5f) : 0;
    const int bx0 = ((int)blockIdx.x * TW + mdx - (RW - (TW + 3)) / 2) & ~3;
    const int by0 = (int)blockIdx.y * TH + mdy - (RH - (TH + 3)) / 2;

    const float *img = in1 + (size_t)b * C * HW;
    float *o = out + (size_t)b * out_bs;

    // out-of-range pixels copy input1 for every channel (:2814-2819)
#pragma unroll
    for (int p = 0; p < PX; ++p)
        if (inside[p] && !active[p])
            for (int c = 0; c < C; ++c) put(o + (size_t)c * HW + pix[p], __ldg(img + (size_t)c * HW + pix[p]));
    if (!any) return;

    // ---- region offsets of the windows (or plane offsets for pixels outside the region), channel-group pipeline ----
    bool use_s[PX];
    int so[PX][4], sc[PX][4];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        const FiPix fp = fi_pixel(w_i, blockIdx.y * TH + ty + (TH / PX) * p, fxy[p][0], fxy[p][1], W, H, 4);
        int ry[4], cx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ry[j] = clampi(fp.T + j, 0, H - 1);   // :2751
            cx[j] = clampi(fp.L + j, 0, W - 1);   // :2753
        }
        use_s[p] = active[p] && cx[0] >= bx0 && cx[3] < bx0 + RW && ry[0] >= by0 && ry[3] < by0 + RH;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            so[p][j] = use_s[p] ? (ry[j] - by0) * RW : ry[j] * W;
            sc[p][j] = use_s[p] ? cx[j] - bx0 : cx[j];
        }
    }
    const int ngroups = (C + CG - 1) / CG;
    auto issue = [&](int g) {   // thread 0: one box [RW, RH, CG] per group; rows / columns / channels outside are zero-filled
        const int s = g % STAGES;
        mbar_arrive_expect_tx(&s_full[s], STAGE_BYTES);
        tma_load_4d(s_img + (size_t)s * CG * PLANE, &map_img, &s_full[s], bx0, by0, g * CG, b);
    };
    if (tid == 0)
        for (int g = 0; g < STAGES && g < ngroups; ++g) issue(g);

    for (int g = 0; g < ngroups; ++g) {
        const int s = g % STAGES;
        mbar_wait_guarded(&s_full[s], (uint32_t)((g / STAGES) & 1));
        const float *reg = s_img + (size_t)s * CG * PLANE;
        const int nc = min(CG, C - g * CG);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            if (!active[p]) continue;
            float *op = o + (size_t)(g * CG) * HW + pix[p];
            // two copies of the loop so that the region reads are LDS (a selected pointer would make them generic loads)
            if (use_s[p]) {
#pragma unroll
                for (int c = 0; c < CG; ++c) {
                    if (c >= nc) break;
                    const float *pl = reg + c * PLANE;
                    float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] =
                                fmaf(pl[so[p][j] + sc[p][i]], w[p][j * 4 + i], Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)]);
                    put(op + (size_t)c * HW, q[p][0] * Q[0] + q[p][1] * Q[1] + q[p][2] * Q[2] + q[p][3] * Q[3]);
#ifdef VFIDKR_BOUNDS_CHECK
                    for (int j = 0; j < 4; ++j)
                        for (int i = 0; i < 4; ++i) {
                            const int idx = so[p][j] + sc[p][i];
                            const bool inside = idx >= 0 && idx < PLANE;
                            const float want = __ldg(img + (size_t)(g * CG + c) * HW + (size_t)(so[p][j] / RW + by0) * W + (sc[p][i] + bx0));
                            bounds_check(inside && __float_as_uint(pl[inside ? idx : 0]) == __float_as_uint(want));
                        }
#endif
                }
            } else {   // pixel far from the tile's mean flow: clamped gathers from the plane
                for (int c = 0; c < nc; ++c) {
                    const float *pl = img + (size_t)(g * CG + c) * HW;
                    float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)] =
                                fmaf(__ldg(pl + so[p][j] + sc[p][i]), w[p][j * 4 + i], Q[(j < 2 ? 0 : 2) + (i < 2 ? 0 : 1)]);
                    put(op + (size_t)c * HW, q[p][0] * Q[0] + q[p][1] * Q[1] + q[p][2] * Q[2] + q[p][3] * Q[3]);
                }
            }
        }
        __syncthreads();   // every thread is done with stage s before it is refilled
        if (tid == 0 && g + STAGES < ngroups) issue(g + STAGES);
    }
}

}  // namespace

// Returns VFIDKR_OK / VFIDKR_ERR_CUDA when the kernel was launched, -1 when it does not apply.
int fi_bigc_forward_ori(const float *in1, const float *in2, const float *in3, float *out,
                        int B, int C, int H, int W, float scale, int accumulate, size_t out_bs, cudaStream_t s)
{
    using namespace bigc;
    if (C <= 4 || W % 4 != 0 || !aligned16(in1) || ceil_div(H, TH) > 65535u) return -1;
    CUtensorMap mimg;
    if (!encode_tensor_map_4d(&mimg, in1, W, H, C, B, RW, RH, CG)) return -1;
    const bool blend = scale != 1.0f || accumulate != 0 || out_bs != (size_t)C * H * W;
    auto kernel = blend ? fi_forward_ori_bigc_kernel<true> : fi_forward_ori_bigc_kernel<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    dim3 grid(ceil_div(W, TW), ceil_div(H, TH), B);
    kernel<<<grid, NT, SMEM_BYTES, s>>>(mimg, in1, in2, in3, out, C, H, W, scale, accumulate, out_bs);
    note_launch();
    return check_launch("filterinterpolation forward (many channels)");
}

}  // namespace vfidkr

VFIDKR_BOUNDS_ACCESSOR(bounds_counts_bigc)
