// correlation_tc.cu -- PWC-Net cost volume (kernel_size 1, strides 1, max_displacement 4, pad 4) on the 5th-generation
// tensor cores of sm_100a: tcgen05.mma kind::tf32 with the accumulator in tensor memory, error-compensated (3 x TF32).
//
// What is computed follows PWCNet/correlation_package_pytorch1_0/correlation_cuda_kernel.cu:74-147:
//   out[b, (tj+4)*9 + (ti+4), y, x] = 1/C * sum_c f1[b,c,y,x] * f2[b,c,y+tj,x+ti],  zero outside the plane.
// As a GEMM: for a PATCH of 8 x 16 output pixels (M = 128 rows) and its 16 x 24 pixel neighbourhood in f2 (N = 384 columns,
// issued as two interleaved halves -- even / odd rows -- of 8 x 24 = 192), D[m][n] = sum_c f1[c, pixel m] * f2[c, position n]; the 81 displacements of a
// pixel are 81 of its row's 384 entries (21 % of the MMA work is useful -- the band structure of the op; a row-tile
// formulation would use 9 of 136).  fp32 parity at 1e-5 needs more than TF32's 10 mantissa bits: each operand is split
// into hi = tf32(v) and lo = tf32(v - hi) and the product is hi*hi + lo*hi + hi*lo (three MMAs, fp32 accumulation in TMEM).
// Tensor rate measured on this part: 2047 MAC/clk/SM for kind::tf32 at M = 128 (tools/microbench/umma_tf32.cu); the SIMT
// kernel of correlation.cu sustains ~53 MAC/clk/SM, so even at 21 % / 3 the tensor pipe has 2.7x the headroom.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  epilogue: TMEM -> registers (tcgen05.ld, a warp owns the 32 TMEM lanes = pixels of its quarter), the
//              lane-dependent band extraction goes through a private shared-memory row, scaled stores to the 81 planes;
//   warps 4-19 stagers: f1 patch / f2 neighbourhood chunk of 32 channels from global memory (L2), hi / lo split, written
//              to shared memory in the canonical K-major no-swizzle core-matrix layout ([K/4][rows/8][8][4] floats, 8-row
//              groups 144 B apart).  TMA cannot do this: the maps are channel-planar, i.e. MN-major operands, and kind::tf32
//              with an MN-major operand returns zeros on sm_100a (tools/microbench/umma_tf32.cu) -- the transposition
//              into K-major rows is what these warps are for, and what limits the kernel on large maps;
//   warp 20    one thread issues the MMAs (12 per stage: 4 K-steps x 3 passes), commits stage-free and accumulator-full.
// A unit of work is (patch, neighbourhood half); the two 192-column accumulators ping-pong in TMEM (512 columns), so the
// epilogue of one unit overlaps the MMAs of the next; operand stages form a 2-deep ring (90 KB each).
#include "common.cuh"
#include "tma.cuh"

namespace vfidkr {
namespace {

namespace ctc {
constexpr int PR = 8, PC = 16, M = PR * PC;                 // output patch
constexpr int NR = 8, NC = PC + 8, NH = NR * NC;            // one neighbourhood half: 8 rows x 24 columns = 192
constexpr int KC = 32, KG = KC / 4;                         // channels per stage, groups of four
constexpr int STAGES = 2;
// shared-memory operand layout (K-major, no swizzle): [K group of 4 channels][8-row group][8 rows][4 floats]; the 8-row
// groups are SBO = 144 bytes apart instead of 128 -- 16 bytes of padding that make the stagers' 16-byte stores (lanes = groups
// of four pixels, 64 bytes apart) conflict-free; K groups are LBO = (rows / 8) * SBO apart
constexpr int SBO = 144, SBO_F = SBO / 4;
constexpr int A_LBO = (M / 8) * SBO, B_LBO = (NH / 8) * SBO;
constexpr int A_FLOATS = KG * A_LBO / 4, B_FLOATS = KG * B_LBO / 4;
constexpr int STAGE_FLOATS = 2 * (A_FLOATS + B_FLOATS);     // hi + lo of both operands
__host__ __device__ constexpr int row_off(int m) { return (m >> 3) * SBO_F + (m & 7) * 4; }   // float offset of row m inside a K group
constexpr int EPI_WARPS = 4, STAGE_WARPS = 16, NTHREADS = (EPI_WARPS + STAGE_WARPS + 1) * 32;
constexpr int EPI_PITCH = 28;                               // 16-byte aligned rows, conflict-free 128-bit stores
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float) + (size_t)M * EPI_PITCH * sizeof(float) + 256;
constexpr uint32_t TMEM_COLS = 512, ACC_STRIDE = 256;       // accumulator b lives at columns [256 b, 256 b + 192)

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // K-major, no swizzle: start address, leading (K-group) byte offset, stride (8-row group) byte offset, sm_100 version bit
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: fp32 accumulator, tf32 x tf32, both operands K-major, N >> 3, M >> 4
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, bool accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)   // arrives on `bar` when every MMA issued so far has completed
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 3 x TF32 split: hi = v with the mantissa cut to TF32's 10 bits (a mask: full-rate, where cvt.rna.tf32 issues at 1/8 of the
// fp32 rate and there are two per operand element), lo = v - hi (exact in fp32; the tensor core reads its top 19 bits).
// |v - hi - tf32(lo)| <= 2^-21 |v|: the dropped lo*lo term and this residual are ~1e-6 of a product, inside the 1e-5 parity bound.
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ void wait_phase(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)     // try_wait suspends the thread for a hardware time slice per probe
        if (spins > (1u << 26)) __trap();      // a protocol error traps instead of hanging the GPU
}
// a whole warp waits: one lane polls, the warp re-converges behind it
__device__ __forceinline__ void warp_wait_phase(uint64_t *bar, uint32_t parity, int lane)
{
    if (lane == 0) wait_phase(bar, parity);
    __syncwarp();
}
}  // namespace ctc

__global__ void __launch_bounds__(ctc::NTHREADS, 1)
corr_forward_tc_kernel(const float *__restrict__ in1, const float *__restrict__ in2, float *__restrict__ out,
                       float *__restrict__ out_b, int bsplit, int C, int H, int W, int tiles_x, int tiles_y, int num_tiles, const FastDiv div_tx, const FastDiv div_tile_img,
                       int vec4)
{
    using namespace ctc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_stage = reinterpret_cast<float *>(smem_raw);                         // [STAGES][A_hi | A_lo | B_hi | B_lo]
    float *s_epi = s_stage + STAGES * STAGE_FLOATS;                               // [M][EPI_PITCH]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_epi + M * EPI_PITCH);
    uint64_t *full = s_bar, *empty = s_bar + STAGES, *acc_full = s_bar + 2 * STAGES, *acc_empty = acc_full + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t HW = (size_t)H * W;
    const int nchunks = (C + KC - 1) / KC;
    const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        // one arrival per WARP (its lane 0, behind a __syncwarp): 256 arrivals on one shared-memory word serialise
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], STAGE_WARPS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == EPI_WARPS + STAGE_WARPS) {       // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    auto decode = [&](int i, int &b, int &y0, int &x0) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        b = div_tile_img.quot(t);
        const int r = t - b * tiles_x * tiles_y, ty = div_tx.quot(r);
        y0 = ty * PR;
        x0 = (r - ty * tiles_x) * PC;
    };

    if (warp >= EPI_WARPS && warp < EPI_WARPS + STAGE_WARPS) {
        // ================================ stagers ================================
        const int st = tid - EPI_WARPS * 32;                         // 0 .. 511
        int it = 0;
        if (vec4) {
            // Rows are whole float4s (W % 4 == 0, aligned bases): a task is (K group g of four channels, four adjacent
            // pixels) = four 128-bit loads, one per channel plane, transposed in registers into four 16-byte rows
            // [pixel][4 channels] of the operand layout.  Patch: 8 x 32 tasks; neighbourhood half: 8 x 48 tasks.
            // 640 tasks per stage over 512 threads: thread st takes task st (patch tasks 0 .. 255, neighbourhood tasks 256 .. 639)
            // and threads 0 .. 127 also task 512 + st.  Loop-invariant geometry first.
            bool isA[2], has[2];
            int gq[2], rq[2], cq[2], offq[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int q = st + 512 * k;
                has[k] = q < 640;
                isA[k] = q < 256;
                const int qb = q - 256;
                gq[k] = isA[k] ? q >> 5 : qb / 48;
                rq[k] = isA[k] ? (q & 31) >> 2 : (qb % 48) / 6;          // patch row / neighbourhood row of this half
                cq[k] = isA[k] ? (q & 3) * 4 : (qb % 6) * 4;             // first of the four columns
                offq[k] = isA[k] ? gq[k] * (A_LBO / 4) + row_off(rq[k] * PC + cq[k]) : gq[k] * (B_LBO / 4) + row_off(rq[k] * NC + cq[k]);
            }
            for (int i = 0; i < my_tiles; ++i) {
                int b, y0, x0;
                decode(i, b, y0, x0);
                // items b >= bsplit are the second problem of a pair launch: the same maps with their roles exchanged
                const bool second = b >= bsplit;
                const size_t boff = (size_t)(second ? b - bsplit : b) * C * HW;
                const float *f1 = (second ? in2 : in1) + boff, *f2 = (second ? in1 : in2) + boff;
                for (int half = 0; half < 2; ++half)
                    for (int ch = 0; ch < nchunks; ++ch, ++it) {
                        const int s = it % STAGES;
                        if (it >= STAGES) warp_wait_phase(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1), lane);   // the MMAs that read this stage are done
                        float *sa_hi = s_stage + s * STAGE_FLOATS, *sa_lo = sa_hi + A_FLOATS, *sb_hi = sa_lo + A_FLOATS, *sb_lo = sb_hi + B_FLOATS;
                        float4 v[2][4];
                        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {      // every load of the stage before the first store
                            const int y = isA[k] ? y0 + rq[k] : y0 - 4 + 2 * rq[k] + half;
                            const int x = isA[k] ? x0 + cq[k] : x0 - 4 + cq[k];
                            const int c0 = ch * KC + 4 * gq[k];
                            const bool in = has[k] && y >= 0 && y < H && x >= 0 && x < W;
                            const float *src = (isA[k] ? f1 : f2) + (size_t)c0 * HW + (in ? (size_t)y * W + x : 0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) v[k][c] = (in && c0 + c < C) ? __ldg(reinterpret_cast<const float4 *>(src + (size_t)c * HW)) : z4;
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (!has[k]) continue;
                            float *hi_base = (isA[k] ? sa_hi : sb_hi) + offq[k], *lo_base = (isA[k] ? sa_lo : sb_lo) + offq[k];
                            const float4 (&t)[4] = v[k];
                            const float px[4][4] = {{t[0].x, t[1].x, t[2].x, t[3].x}, {t[0].y, t[1].y, t[2].y, t[3].y},
                                                    {t[0].z, t[1].z, t[2].z, t[3].z}, {t[0].w, t[1].w, t[2].w, t[3].w}};   // [pixel][channel]
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 hi, lo;
                                hi.x = tf32_hi(px[j][0]); hi.y = tf32_hi(px[j][1]); hi.z = tf32_hi(px[j][2]); hi.w = tf32_hi(px[j][3]);
                                lo.x = px[j][0] - hi.x; lo.y = px[j][1] - hi.y; lo.z = px[j][2] - hi.z; lo.w = px[j][3] - hi.w;
                                *reinterpret_cast<float4 *>(hi_base + 4 * j) = hi;      // rows m .. m + 3 of one 8-row group: 16 bytes apart
                                *reinterpret_cast<float4 *>(lo_base + 4 * j) = lo;
                            }
                        }
                        fence_proxy_async();             // this thread's operand stores -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                    }
            }
        } else {
            // any width / alignment: a task is (K group, one pixel), four scalar loads
            constexpr int A_TASKS = KG * M, TASKS = KG * (M + NH), PER = TASKS / (STAGE_WARPS * 32);   // 2560 tasks, 5 per thread
            static_assert(TASKS % (STAGE_WARPS * 32) == 0, "tasks per stager thread");
            for (int i = 0; i < my_tiles; ++i) {
                int b, y0, x0;
                decode(i, b, y0, x0);
                // items b >= bsplit are the second problem of a pair launch: the same maps with their roles exchanged
                const bool second = b >= bsplit;
                const size_t boff = (size_t)(second ? b - bsplit : b) * C * HW;
                const float *f1 = (second ? in2 : in1) + boff, *f2 = (second ? in1 : in2) + boff;
                for (int half = 0; half < 2; ++half)
                    for (int ch = 0; ch < nchunks; ++ch, ++it) {
                        const int s = it % STAGES;
                        if (it >= STAGES) warp_wait_phase(&empty[s], (uint32_t)(((it / STAGES) - 1) & 1), lane);
                        float *sa_hi = s_stage + s * STAGE_FLOATS, *sa_lo = sa_hi + A_FLOATS, *sb_hi = sa_lo + A_FLOATS, *sb_lo = sb_hi + B_FLOATS;
                        float v[PER][4];
#pragma unroll
                        for (int k = 0; k < PER; ++k) {      // every load of the stage before the first store
                            const int q = st + k * (STAGE_WARPS * 32);
                            const bool isA = k < A_TASKS / (STAGE_WARPS * 32);      // compile-time per k: tasks 0 .. 1023 are the patch
                            const int qq = isA ? q : q - A_TASKS;
                            const int g = isA ? qq / M : qq / NH, m = isA ? qq % M : qq % NH;
                            const int y = isA ? y0 + m / PC : y0 - 4 + 2 * (m / NC) + half;
                            const int x = isA ? x0 + m % PC : x0 - 4 + m % NC;
                            const bool in = y >= 0 && y < H && x >= 0 && x < W;
                            const float *src = (isA ? f1 : f2) + (size_t)(ch * KC + 4 * g) * HW + (in ? (size_t)y * W + x : 0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) v[k][c] = (in && ch * KC + 4 * g + c < C) ? __ldg(src + (size_t)c * HW) : 0.0f;
                        }
#pragma unroll
                        for (int k = 0; k < PER; ++k) {
                            const int q = st + k * (STAGE_WARPS * 32);
                            const bool isA = k < A_TASKS / (STAGE_WARPS * 32);
                            const int qq = isA ? q : q - A_TASKS;
                            const int g = isA ? qq / M : qq / NH, m = isA ? qq % M : qq % NH;
                            const int off = g * ((isA ? A_LBO : B_LBO) / 4) + row_off(m);
                            float4 hi, lo;
                            hi.x = tf32_hi(v[k][0]); hi.y = tf32_hi(v[k][1]); hi.z = tf32_hi(v[k][2]); hi.w = tf32_hi(v[k][3]);
                            lo.x = v[k][0] - hi.x; lo.y = v[k][1] - hi.y; lo.z = v[k][2] - hi.z; lo.w = v[k][3] - hi.w;
                            *reinterpret_cast<float4 *>((isA ? sa_hi : sb_hi) + off) = hi;
                            *reinterpret_cast<float4 *>((isA ? sa_lo : sb_lo) + off) = lo;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                    }
            }
        }
    } else if (warp == EPI_WARPS + STAGE_WARPS) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            int it = 0, u = 0;
            for (int i = 0; i < my_tiles; ++i)
                for (int half = 0; half < 2; ++half, ++u) {
                    const int ab = u & 1;
                    if (u >= 2) wait_phase(&acc_empty[ab], (uint32_t)(((u >> 1) - 1) & 1));      // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d = tmem + ab * ACC_STRIDE;
                    for (int ch = 0; ch < nchunks; ++ch, ++it) {
                        const int s = it % STAGES;
                        wait_phase(&full[s], (uint32_t)((it / STAGES) & 1));
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(s_stage + s * STAGE_FLOATS), a_lo = a_hi + A_FLOATS * 4;
                        const uint32_t b_hi = a_lo + A_FLOATS * 4, b_lo = b_hi + B_FLOATS * 4;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t oa = ks * 2 * A_LBO, ob = ks * 2 * B_LBO;       // two K groups per MMA (K = 8)
                            const uint64_t dah = smem_desc(a_hi + oa, A_LBO, SBO), dal = smem_desc(a_lo + oa, A_LBO, SBO);
                            const uint64_t dbh = smem_desc(b_hi + ob, B_LBO, SBO), dbl = smem_desc(b_lo + ob, B_LBO, SBO);
                            umma_tf32(d, dah, dbh, ch > 0 || ks > 0);
                            umma_tf32(d, dal, dbh, true);
                            umma_tf32(d, dah, dbl, true);
                        }
                        umma_commit(&empty[s]);                 // stage s is free once these MMAs have read it
                    }
                    umma_commit(&acc_full[ab]);                 // the unit's accumulator is complete
                }
        }
    } else {
        // ================================ epilogue ================================
        const int m = tid;                                      // TMEM lane = patch pixel
        const int r = m / PC, col = m % PC;
        float *row = s_epi + m * EPI_PITCH;
        const float inv = 1.0f / (float)C;                      // nelems = kernel_size^2 * C (:104); one reciprocal, as in correlation.cu
        const int r_lo = 2 * warp;                              // the warp's two patch rows: 2 warp, 2 warp + 1
        int u = 0;
        for (int i = 0; i < my_tiles; ++i) {
            int b, y0, x0;
            decode(i, b, y0, x0);
            const int y = y0 + r, x = x0 + col;
            const bool live = y < H && x < W;
            float *o = (b >= bsplit ? out_b + (size_t)(b - bsplit) * 81 * HW : out + (size_t)b * 81 * HW) + (size_t)y * W + x;
            for (int half = 0; half < 2; ++half, ++u) {
                const int ab = u & 1;
                warp_wait_phase(&acc_full[ab], (uint32_t)((u >> 1) & 1), lane);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ab * ACC_STRIDE;
#pragma unroll 1
                for (int nr = 0; nr < NR; ++nr) {
                    // row nr of this half is neighbourhood row na = 2 nr + half (the halves INTERLEAVE the 16 rows, so that every
                    // warp needs five rows of each -- with the rows split 0-7 / 8-15 the warps needed 8, 6, 4, 2 of one half and
                    // the unit took as long as the busiest warp); it is displacement tj = na - r - 4 of a pixel in patch row r.
                    // Rows no pixel of this warp needs are skipped (warp-uniform).
                    const int na = 2 * nr + half;
                    if (na < r_lo || na > r_lo + 9) continue;
                    uint32_t a[8], bq[8], c[8];
                    tmem_ld8(taddr + nr * NC, a);
                    tmem_ld8(taddr + nr * NC + 8, bq);
                    tmem_ld8(taddr + nr * NC + 16, c);
                    tmem_wait_ld();
                    float4 *row4 = reinterpret_cast<float4 *>(row);
                    row4[0] = make_float4(__uint_as_float(a[0]), __uint_as_float(a[1]), __uint_as_float(a[2]), __uint_as_float(a[3]));
                    row4[1] = make_float4(__uint_as_float(a[4]), __uint_as_float(a[5]), __uint_as_float(a[6]), __uint_as_float(a[7]));
                    row4[2] = make_float4(__uint_as_float(bq[0]), __uint_as_float(bq[1]), __uint_as_float(bq[2]), __uint_as_float(bq[3]));
                    row4[3] = make_float4(__uint_as_float(bq[4]), __uint_as_float(bq[5]), __uint_as_float(bq[6]), __uint_as_float(bq[7]));
                    row4[4] = make_float4(__uint_as_float(c[0]), __uint_as_float(c[1]), __uint_as_float(c[2]), __uint_as_float(c[3]));
                    row4[5] = make_float4(__uint_as_float(c[4]), __uint_as_float(c[5]), __uint_as_float(c[6]), __uint_as_float(c[7]));
                    __syncwarp();
                    const int tj = na - r - 4;
                    if (live && tj >= -4 && tj <= 4) {
                        float *ot = o + (size_t)((tj + 4) * 9) * HW;
#pragma unroll
                        for (int ti = 0; ti < 9; ++ti) st_stream(ot + (size_t)ti * HW, row[col + ti] * inv);   // columns x - 4 + ti
                    }
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ab]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + STAGE_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace

// Returns VFIDKR_OK / VFIDKR_ERR_CUDA when the kernel was launched, -1 when it does not apply.
// out_b != nullptr: a pair launch, out_b = correlation(in2, in1) next to out = correlation(in1, in2).
int corr_forward_tc(const float *in1, const float *in2, float *out, float *out_b, int B, int C, int H, int W, cudaStream_t s)
{
    using namespace ctc;
    const int tiles_x = ceil_div(W, PC), tiles_y = ceil_div(H, PR);
    const long long num_tiles = (long long)tiles_x * tiles_y * B * (out_b ? 2 : 1);
    if (num_tiles >= (1ll << 30) || (long long)C * H * W >= (1ll << 31)) return -1;
    if (cudaFuncSetAttribute(corr_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    const int nblk = (int)std::min<long long>(num_tiles, (long long)sm_count());
    const int vec4 = (W % 4 == 0 && aligned16(in1) && aligned16(in2)) ? 1 : 0;
    corr_forward_tc_kernel<<<nblk, NTHREADS, SMEM_BYTES, s>>>(in1, in2, out, out_b, out_b ? B : 2 * B, C, H, W, tiles_x, tiles_y, (int)num_tiles,
                                                              FastDiv((unsigned)tiles_x), FastDiv((unsigned)(tiles_x * tiles_y)), vec4);
    note_launch();
    return check_launch("correlation forward (tensor cores)");
}

}  // namespace vfidkr
