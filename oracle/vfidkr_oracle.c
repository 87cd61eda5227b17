/*
 * vfidkr_oracle.c -- CPU restatement of the VFIDKR per-pixel sampling / warping kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import, link or execute this
 * file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: PINNED against the reference itself.  The reference ships no golden vectors, no
 * known-answer tests and no CPU implementation for this path, but its CUDA
 * extensions compile unmodified for sm_100a (oracle/build_ref.py -> oracle/_ref/*.so); their outputs on
 * the seeded cases of tests/golden_cases.py are committed as tests/golden/*.npz (oracle/make_golden.py)
 * and tests/test_golden.py::test_oracle_reproduces_reference_kernels holds this file to them.  On top:
 * the known-answer identities of SURVEY.md section 8c (tests/test_oracle_kat.py) and finite-difference
 * gradient checks of the float64 forward.
 *
 * Conventions:
 *   - every input is float32, NCHW, contiguous -- the same bytes the GPU sees;
 *   - index / offset / predicate arithmetic is done in float32 exactly as the CUDA source
 *     writes it (x2 = (float)w + fx, truncating (int) casts, fp32 comparisons);
 *   - value accumulation is float64, outputs are float64;
 *   - scatter ("atomicAdd") targets accumulate in float64 in a fixed order, so the oracle
 *     is deterministic.
 *   - reads of the deformable variants whose row/column index leaves the channel plane are
 *     undefined behaviour in the reference (hard part 1).  The oracle
 *     (and the CUDA product) define them by clamping the READ index to the plane while
 *     keeping the reference's weights; inside the in-contract domain nothing changes.
 *
 * All file:line citations are relative to /root/reference/.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return imin(imax(v, lo), hi); }
/* float -> int truncation toward zero.  In C the conversion is undefined outside the int range; CUDA's
 * cvt.rzi.s32.f32 saturates and maps NaN to 0.  The oracle adopts the CUDA definition so that
 * out-of-contract offsets (DKR families) stay defined. */
static inline int f2i_rz(float v)
{
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* ------------------------------------------------------------------------------------------
 * FilterInterpolation -- four kernel families behind one restatement.
 *   variant 0  "_ori"                    my_package/FilterInterpolation/filterinterpolation_cuda_kernel.cu:2692-3125
 *   variant 1  4-input DKR (static quad)  :29-426 (fwd), :430-1215 (bwd)
 *   variant 2  "_deforconv" (dynamic quad) :1353-1496 (fwd), :1500-1935 (bwd)
 *   variant 3  "_nofilterwithdeforconv"    :2070-2191 (fwd), :2195-2567 (bwd); offsets arrive in input3
 * ------------------------------------------------------------------------------------------ */
enum { FI_ORI = 0, FI_DKR = 1, FI_DEFORCONV = 2, FI_NOFILTER = 3 };

typedef struct {
    int in_range;      /* range test :2735-2736 */
    int ix, iy;        /* (int)x2, (int)y2 */
    int L, T;          /* ix2_L, iy2_T :2737-2738 */
    float x2, y2;
    float alpha, beta; /* :2742-2743 */
} fi_pixel_t;

static fi_pixel_t fi_pixel(int w_i, int h_i, float fx, float fy, int w, int h, int F)
{
    fi_pixel_t p;
    /* :2731-2732  x2 = (float)(w_i) + fx  -- float32 */
    volatile float x2 = (float)w_i + fx;
    volatile float y2 = (float)h_i + fy;
    p.x2 = x2; p.y2 = y2;
    /* :2735-2736 */
    p.in_range = (x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(w - 1) && y2 <= (float)(h - 1)
                  && fabsf(fx) < (float)w / 2.0f && fabsf(fy) < (float)h / 2.0f);
    p.ix = p.iy = p.L = p.T = 0; p.alpha = p.beta = 0.0f;
    if (p.in_range) {
        p.ix = (int)x2; p.iy = (int)y2;
        p.L = p.ix + 1 - (int)(F / 2);
        p.T = p.iy + 1 - (int)(F / 2);
        volatile float a = x2 - (float)p.ix;  /* float32 subtraction :2742 */
        volatile float b = y2 - (float)p.iy;
        p.alpha = a; p.beta = b;
    }
    return p;
}

/* Deformed tap: bilinear sample of one channel plane at (cy + offY, cx + offX).
 * Restates :98-112 (identical text in every DKR family).  Returns the sample and its
 * derivatives w.r.t. the y- and x- offsets (:986-989 and :1104-1107). */
typedef struct {
    float fracX, fracY;
    int   Top, Left, Bottom, Right;       /* unclamped, as the reference computes them */
    float phiX, phiY;
} fi_deform_t;

static fi_deform_t fi_deform(int cy, int cx, float offY, float offX)
{
    fi_deform_t d;
    volatile float fy = (float)cy + offY;   /* float fracY = _filter_j + input4[...]  :98 */
    volatile float fx = (float)cx + offX;   /* :99 */
    d.fracY = fy; d.fracX = fx;
    d.Top = f2i_rz(fy); d.Left = f2i_rz(fx); /* :102-103 truncation toward zero */
    volatile float py = fy - (float)d.Top;  /* :100 */
    volatile float px = fx - (float)d.Left; /* :101 */
    d.phiY = py; d.phiX = px;
    /* :104-105, NOT clamped in the reference (+1 guarded against int overflow) */
    d.Bottom = d.Top < 2147483647 ? d.Top + 1 : d.Top; d.Right = d.Left < 2147483647 ? d.Left + 1 : d.Left;
    return d;
}

static inline void fi_sample(const float *plane, int h, int w, const fi_deform_t *d,
                             double *S, double *dSy, double *dSx)
{
    /* read indices clamped to the plane (memory-safety rule, see header) */
    int t = iclamp(d->Top, 0, h - 1), b = iclamp(d->Bottom, 0, h - 1);
    int l = iclamp(d->Left, 0, w - 1), r = iclamp(d->Right, 0, w - 1);
    double vTL = plane[(size_t)t * w + l], vTR = plane[(size_t)t * w + r];
    double vBL = plane[(size_t)b * w + l], vBR = plane[(size_t)b * w + r];
    /* the P-weights are formed in float32 in the reference (:106-109); they are values, so
     * the oracle forms them in float64 from the float32 phi's */
    double px = d->phiX, py = d->phiY;
    double PTL = (1 - px) * (1 - py), PTR = px * (1 - py), PBL = (1 - px) * py, PBR = py * px;
    *S = PTL * vTL + PTR * vTR + PBL * vBL + PBR * vBR;              /* :110-111 */
    *dSy = -(1 - px) * vTL + (1 - px) * vBL - px * vTR + px * vBR;   /* :986-989 */
    *dSx = -(1 - py) * vTL + (1 - py) * vTR - py * vBL + py * vBR;   /* :1104-1107 */
}

/* quadrant index of tap (j,i): 0=TL 1=TR 2=BL 3=BR */
static inline int fi_quadrant(int variant, const fi_pixel_t *p, int j, int i, int F,
                              const fi_deform_t *d)
{
    int top, left;
    if (variant == FI_ORI || variant == FI_DKR) {
        /* static split: rows filter_j <= (int)y2, columns filter_i <= (int)x2 (:2750-2787, :91-198) */
        top = (p->T + j) <= p->iy;
        left = (p->L + i) <= p->ix;
    } else {
        /* data-dependent split (:1442-1468, :2135-2151): fracX <= x2, fracY <= y2 in float32 */
        top = d->fracY <= p->y2;
        left = d->fracX <= p->x2;
    }
    (void)F;
    return (top ? 0 : 2) + (left ? 0 : 1);
}

/* Forward for all four families.
 *   input1 [B,C,H,W]  input2 [B,2,H,W]  input3 [B,F*F,H,W] (variant 3: [B,2*F*F,H,W] offsets)
 *   input4 [B,2*F*F,H,W] (variants 1,2; ignored otherwise)   output [B,C,H,W] float64
 * Returns 0, or 1 on the argument errors the .cc glue rejects (filterinterpolation_cuda.cc:24-60). */
ORACLE_API int oracle_fi_forward(int variant, const float *input1, const float *input2,
                                 const float *input3, const float *input4, double *output,
                                 int B, int C, int H, int W, int F)
{
    if (variant < 0 || variant > 3 || F <= 0) return 1;
    const size_t HW = (size_t)H * W;
    const int T2 = F * F;
    const float *offs = (variant == FI_NOFILTER) ? input3 : input4;
    const size_t offs_b = (size_t)2 * T2 * HW;
    const size_t filt_b = (variant == FI_NOFILTER) ? offs_b : (size_t)T2 * HW;
    /* the 4-input forward only runs for filter_size 4 or 6 (:68); otherwise the zero-filled
     * output is left untouched */
    if (variant == FI_DKR && !(F == 4 || F == 6)) {
        memset(output, 0, sizeof(double) * (size_t)B * C * HW);
        return 0;
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input2[((size_t)b * 2 + 0) * HW + pix];
                float fy = input2[((size_t)b * 2 + 1) * HW + pix];
                fi_pixel_t p = fi_pixel(w_i, h_i, fx, fy, W, H, F);
                if (!p.in_range) {
                    /* :2814-2819 -- copies input1 (despite the "fill zeros" comment) */
                    for (int c = 0; c < C; ++c)
                        output[((size_t)b * C + c) * HW + pix] = input1[((size_t)b * C + c) * HW + pix];
                    continue;
                }
                double a = p.alpha, be = p.beta;
                double q[4] = { (1 - a) * (1 - be), a * (1 - be), (1 - a) * be, a * be }; /* :2789-2793 */
                for (int c = 0; c < C; ++c) {
                    const float *plane = input1 + ((size_t)b * C + c) * HW;
                    double Q[4] = { 0, 0, 0, 0 };
                    for (int j = 0; j < F; ++j) {
                        int cy = iclamp(p.T + j, 0, H - 1);          /* :2751 */
                        for (int i = 0; i < F; ++i) {
                            int cx = iclamp(p.L + i, 0, W - 1);      /* :2753 */
                            int k = j * F + i;
                            double wgt = (variant == FI_NOFILTER) ? 1.0
                                        : (double)input3[(size_t)b * filt_b + (size_t)k * HW + pix];
                            double S; fi_deform_t d; memset(&d, 0, sizeof d);
                            if (variant == FI_ORI) {
                                S = plane[(size_t)cy * W + cx];       /* :2754 */
                            } else {
                                float oy = offs[(size_t)b * offs_b + (size_t)k * HW + pix];
                                float ox = offs[(size_t)b * offs_b + (size_t)(T2 + k) * HW + pix];
                                d = fi_deform(cy, cx, oy, ox);
                                double dy, dx; fi_sample(plane, H, W, &d, &S, &dy, &dx);
                            }
                            Q[fi_quadrant(variant, &p, j, i, F, &d)] += S * wgt;
                        }
                    }
                    output[((size_t)b * C + c) * HW + pix] = q[0] * Q[0] + q[1] * Q[1] + q[2] * Q[2] + q[3] * Q[3];
                }
            }
    return 0;
}

/* Backward for all four families.  All gradient buffers are float64 and are fully written
 * (zeroed here; the reference relies on the caller's zero-fill, FilterInterpolationLayer.py:62-64).
 *   gi1 [B,C,H,W]  gi2 [B,2,H,W]  gi3 like input3  gi4 like input4 (variants 1,2; may be NULL otherwise)
 * Quirks restated on purpose (hard part 2):
 *   - gi1 is scattered to the UNDEFORMED clamped tap (:497-499, :1581-1583, :2258);
 *   - out-of-range pixels contribute nothing although forward copied input1 there;
 *   - variant 3 scatters gi1 unweighted. */
ORACLE_API int oracle_fi_backward(int variant, const float *input1, const float *input2,
                                  const float *input3, const float *input4, const float *gradoutput,
                                  double *gi1, double *gi2, double *gi3, double *gi4,
                                  int B, int C, int H, int W, int F)
{
    if (variant < 0 || variant > 3 || F <= 0) return 1;
    const size_t HW = (size_t)H * W;
    const int T2 = F * F;
    const float *offs = (variant == FI_NOFILTER) ? input3 : input4;
    const size_t offs_b = (size_t)2 * T2 * HW;
    const size_t filt_b = (variant == FI_NOFILTER) ? offs_b : (size_t)T2 * HW;
    double *goffs = (variant == FI_NOFILTER) ? gi3 : gi4;
    memset(gi1, 0, sizeof(double) * (size_t)B * C * HW);
    memset(gi2, 0, sizeof(double) * (size_t)B * 2 * HW);
    memset(gi3, 0, sizeof(double) * (size_t)B * filt_b);
    if ((variant == FI_DKR || variant == FI_DEFORCONV) && gi4)
        memset(gi4, 0, sizeof(double) * (size_t)B * offs_b);
    /* parallel over batch items only: the gi1 scatter stays race-free and ordered */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input2[((size_t)b * 2 + 0) * HW + pix];
                float fy = input2[((size_t)b * 2 + 1) * HW + pix];
                fi_pixel_t p = fi_pixel(w_i, h_i, fx, fy, W, H, F);
                if (!p.in_range) continue;
                double a = p.alpha, be = p.beta;
                double q[4] = { (1 - a) * (1 - be), a * (1 - be), (1 - a) * be, a * be };
                double gx = 0.0, gy = 0.0;
                for (int c = 0; c < C; ++c) {
                    const float *plane = input1 + ((size_t)b * C + c) * HW;
                    double *gplane = gi1 + ((size_t)b * C + c) * HW;
                    double g = gradoutput[((size_t)b * C + c) * HW + pix];   /* :2883 */
                    double Q[4] = { 0, 0, 0, 0 };
                    for (int j = 0; j < F; ++j) {
                        int cy = iclamp(p.T + j, 0, H - 1);
                        for (int i = 0; i < F; ++i) {
                            int cx = iclamp(p.L + i, 0, W - 1);
                            int k = j * F + i;
                            double wgt = (variant == FI_NOFILTER) ? 1.0
                                        : (double)input3[(size_t)b * filt_b + (size_t)k * HW + pix];
                            double S, dSy = 0, dSx = 0; fi_deform_t d; memset(&d, 0, sizeof d);
                            if (variant == FI_ORI) {
                                S = plane[(size_t)cy * W + cx];
                            } else {
                                float oy = offs[(size_t)b * offs_b + (size_t)k * HW + pix];
                                float ox = offs[(size_t)b * offs_b + (size_t)(T2 + k) * HW + pix];
                                d = fi_deform(cy, cx, oy, ox);
                                fi_sample(plane, H, W, &d, &S, &dSy, &dSx);
                            }
                            int qi = fi_quadrant(variant, &p, j, i, F, &d);
                            double gq = g * q[qi];                    /* TL_grad etc. :2885 */
                            /* Step 1: image gradient, scattered to the undeformed tap (:2890-2892, :497-499, :2258) */
                            gplane[(size_t)cy * W + cx] += gq * wgt;
                            /* Step 3: filter gradient (:2893-2895, :520-522); not present in variant 3 */
                            if (variant != FI_NOFILTER)
                                gi3[(size_t)b * filt_b + (size_t)k * HW + pix] += gq * S;
                            /* Step 4: offset-field gradient (:990-993 y, :1108-1111 x; variant 3 :2460-2567) */
                            if (variant != FI_ORI) {
                                goffs[(size_t)b * offs_b + (size_t)k * HW + pix] += gq * dSy * wgt;
                                goffs[(size_t)b * offs_b + (size_t)(T2 + k) * HW + pix] += gq * dSx * wgt;
                            }
                            Q[qi] += S * wgt;
                        }
                    }
                    /* Step 2: flow gradient (:2965-3031, :3036-3102) */
                    double gamma = 1.0 - be;  /* :2965 */
                    gx += g * (gamma * (Q[1] - Q[0]) + (1 - gamma) * (Q[3] - Q[2]));
                    gamma = 1.0 - a;          /* :3036 */
                    gy += g * (gamma * (Q[2] - Q[0]) + (1 - gamma) * (Q[3] - Q[1]));
                }
                gi2[((size_t)b * 2 + 0) * HW + pix] = gx;   /* :3031 */
                gi2[((size_t)b * 2 + 1) * HW + pix] = gy;   /* :3102 */
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FlowProjection / DepthFlowProjection
 *   my_package/FlowProjection/flowprojection_cuda_kernel.cu:29-301
 *   my_package/DepthFlowProjection/depthflowprojection_cuda_kernel.cu:29-341
 * depth == NULL selects FlowProjection (weight 1).
 * ------------------------------------------------------------------------------------------ */
ORACLE_API int oracle_flowprojection_forward(const float *input1, const float *depth,
                                             double *count, double *output,
                                             int B, int H, int W, int fillhole)
{
    const size_t HW = (size_t)H * W;
    memset(count, 0, sizeof(double) * (size_t)B * HW);
    memset(output, 0, sizeof(double) * (size_t)B * 2 * HW);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        double *ou = output + ((size_t)b * 2 + 0) * HW, *ov = output + ((size_t)b * 2 + 1) * HW;
        double *cn = count + (size_t)b * HW;
        /* splat: flowprojection_cuda_kernel.cu:59-90, depthflowprojection_cuda_kernel.cu:61-93 */
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input1[((size_t)b * 2 + 0) * HW + pix];
                float fy = input1[((size_t)b * 2 + 1) * HW + pix];
                volatile float x2 = (float)w_i + fx;
                volatile float y2 = (float)h_i + fy;
                if (x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1)) {
                    int L = (int)x2, T = (int)y2;
                    int R = imin(L + 1, W - 1), Bm = imin(T + 1, H - 1);
                    double d = depth ? (double)depth[(size_t)b * HW + pix] : 1.0;
                    const size_t corner[4] = { (size_t)T * W + L, (size_t)T * W + R,
                                               (size_t)Bm * W + L, (size_t)Bm * W + R };
                    for (int k = 0; k < 4; ++k) {   /* a clamped corner is hit twice, as in the reference */
                        ou[corner[k]] += -d * fx;
                        ov[corner[k]] += -d * fy;
                        cn[corner[k]] += d;
                    }
                }
            }
        /* averaging: flowprojection_cuda_kernel.cu:128-135 */
        for (size_t pix = 0; pix < HW; ++pix)
            if (cn[pix] > 0.0) { ou[pix] /= cn[pix]; ov[pix] /= cn[pix]; }
        /* hole filling: flowprojection_cuda_kernel.cu:171-232 (reads only non-hole pixels, which it never writes) */
        if (fillhole) {
            for (int h_i = 0; h_i < H; ++h_i)
                for (int w_i = 0; w_i < W; ++w_i) {
                    const size_t pix = (size_t)h_i * W + w_i;
                    if (cn[pix] > 0.0) continue;
                    int lo = w_i; double lt = 0.0;
                    while (lt == 0.0 && lo - 1 >= 0) { lo--; lt = cn[(size_t)h_i * W + lo]; }
                    int ro = w_i; double rt = 0.0;
                    while (rt == 0.0 && ro + 1 <= W - 1) { ro++; rt = cn[(size_t)h_i * W + ro]; }
                    int uo = h_i; double ut = 0.0;
                    while (ut == 0.0 && uo - 1 >= 0) { uo--; ut = cn[(size_t)uo * W + w_i]; }
                    int dn = h_i; double dt = 0.0;
                    while (dt == 0.0 && dn + 1 <= H - 1) { dn++; dt = cn[(size_t)dn * W + w_i]; }
                    if (lt + rt + ut + dt <= 0.0) continue;
                    double l = lt > 0.0, r = rt > 0.0, u = ut > 0.0, d = dt > 0.0;
                    ou[pix] = (l * ou[(size_t)h_i * W + lo] + r * ou[(size_t)h_i * W + ro] +
                               u * ou[(size_t)uo * W + w_i] + d * ou[(size_t)dn * W + w_i]) / (l + r + u + d);
                    ov[pix] = (l * ov[(size_t)h_i * W + lo] + r * ov[(size_t)h_i * W + ro] +
                               u * ov[(size_t)uo * W + w_i] + d * ov[(size_t)dn * W + w_i]) / (l + r + u + d);
                }
        }
    }
    return 0;
}

/* Backward (gather).  count / output are the float32 tensors the forward saved
 * (FlowProjectionLayer.py:48, DepthFlowProjectionLayer.py:62).
 * flowprojection_cuda_kernel.cu:266-297; depthflowprojection_cuda_kernel.cu:276-337.
 * gi2 (grad wrt depth) and `output` are only used when depth != NULL. */
ORACLE_API int oracle_flowprojection_backward(const float *input1, const float *depth,
                                              const float *count, const float *output,
                                              const float *gradoutput, double *gi1, double *gi2,
                                              int B, int H, int W)
{
    const size_t HW = (size_t)H * W;
    memset(gi1, 0, sizeof(double) * (size_t)B * 2 * HW);
    if (depth && gi2) memset(gi2, 0, sizeof(double) * (size_t)B * HW);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input1[((size_t)b * 2 + 0) * HW + pix];
                float fy = input1[((size_t)b * 2 + 1) * HW + pix];
                volatile float x2 = (float)w_i + fx;
                volatile float y2 = (float)h_i + fy;
                if (!(x2 >= 0.0f && y2 >= 0.0f && x2 <= (float)(W - 1) && y2 <= (float)(H - 1))) continue;
                int L = (int)x2, T = (int)y2;
                int R = imin(L + 1, W - 1), Bm = imin(T + 1, H - 1);
                const size_t corner[4] = { (size_t)T * W + L, (size_t)T * W + R,
                                           (size_t)Bm * W + L, (size_t)Bm * W + R };
                const float *gu = gradoutput + ((size_t)b * 2 + 0) * HW, *gv = gradoutput + ((size_t)b * 2 + 1) * HW;
                const float *cn = count + (size_t)b * HW;
                double d = depth ? (double)depth[(size_t)b * HW + pix] : 1.0;
                double su = 0, sv = 0, sd = 0;
                for (int k = 0; k < 4; ++k) {
                    double cnt = cn[corner[k]];
                    su += -(double)gu[corner[k]] * d / cnt;
                    sv += -(double)gv[corner[k]] * d / cnt;
                    if (depth) {
                        const float *ou = output + ((size_t)b * 2 + 0) * HW, *ov = output + ((size_t)b * 2 + 1) * HW;
                        sd += -(double)gu[corner[k]] / cnt * ((double)fx - (double)ou[corner[k]]);
                        sd += -(double)gv[corner[k]] / cnt * ((double)fy - (double)ov[corner[k]]);
                    }
                }
                gi1[((size_t)b * 2 + 0) * HW + pix] = su;
                gi1[((size_t)b * 2 + 1) * HW + pix] = sv;
                if (depth && gi2) gi2[(size_t)b * HW + pix] = sd;
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Interpolation / InterpolationCh (identical kernels)
 *   my_package/Interpolation/interpolation_cuda_kernel.cu:29-98 (fwd), :102-204 (bwd)
 * ------------------------------------------------------------------------------------------ */
ORACLE_API int oracle_interpolation_forward(const float *input1, const float *input2, double *output,
                                            int B, int C, int H, int W)
{
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input2[((size_t)b * 2 + 0) * HW + pix];
                float fy = input2[((size_t)b * 2 + 1) * HW + pix];
                volatile float x2 = (float)w_i + fx;
                volatile float y2 = (float)h_i + fy;
                if (x2 >= 0.0f && y2 >= 0.0f && x2 < (float)W && y2 < (float)H) {  /* strict :71 */
                    int L = (int)x2, T = (int)y2;
                    int R = imin(L + 1, W - 1), Bm = imin(T + 1, H - 1);
                    volatile float af = x2 - (float)L, bf = y2 - (float)T;        /* :77-78 */
                    double a = af, be = bf;
                    for (int c = 0; c < C; ++c) {
                        const float *pl = input1 + ((size_t)b * C + c) * HW;
                        double TL = pl[(size_t)T * W + L], TR = pl[(size_t)T * W + R];
                        double BL = pl[(size_t)Bm * W + L], BR = pl[(size_t)Bm * W + R];
                        output[((size_t)b * C + c) * HW + pix] =
                            (1 - a) * (1 - be) * TL + a * (1 - be) * TR + (1 - a) * be * BL + a * be * BR; /* :85-86 */
                    }
                } else {
                    for (int c = 0; c < C; ++c) output[((size_t)b * C + c) * HW + pix] = 0.0;  /* :88-92 */
                }
            }
    return 0;
}

ORACLE_API int oracle_interpolation_backward(const float *input1, const float *input2,
                                             const float *gradoutput, double *gi1, double *gi2,
                                             int B, int C, int H, int W)
{
    const size_t HW = (size_t)H * W;
    memset(gi1, 0, sizeof(double) * (size_t)B * C * HW);
    memset(gi2, 0, sizeof(double) * (size_t)B * 2 * HW);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < H; ++h_i)
            for (int w_i = 0; w_i < W; ++w_i) {
                const size_t pix = (size_t)h_i * W + w_i;
                float fx = input2[((size_t)b * 2 + 0) * HW + pix];
                float fy = input2[((size_t)b * 2 + 1) * HW + pix];
                volatile float x2 = (float)w_i + fx;
                volatile float y2 = (float)h_i + fy;
                if (!(x2 >= 0.0f && y2 >= 0.0f && x2 < (float)W && y2 < (float)H)) continue;
                int L = (int)x2, T = (int)y2;
                int R = imin(L + 1, W - 1), Bm = imin(T + 1, H - 1);
                volatile float af = x2 - (float)L, bf = y2 - (float)T;
                double a = af, be = bf;
                /* gamma uses the CLAMPED corner (:163 "iy2_B - y2", :180 "ix2_R - x2"), float32 */
                volatile float g1f = (float)Bm - y2, g2f = (float)R - x2;
                double gam1 = g1f, gam2 = g2f;
                double bx = 0, by = 0;
                for (int c = 0; c < C; ++c) {
                    const float *pl = input1 + ((size_t)b * C + c) * HW;
                    double *gp = gi1 + ((size_t)b * C + c) * HW;
                    double g = gradoutput[((size_t)b * C + c) * HW + pix];
                    gp[(size_t)T * W + L] += g * (1 - a) * (1 - be);   /* :156-159 */
                    gp[(size_t)T * W + R] += g * a * (1 - be);
                    gp[(size_t)Bm * W + L] += g * (1 - a) * be;
                    gp[(size_t)Bm * W + R] += g * a * be;
                    double TL = pl[(size_t)T * W + L], TR = pl[(size_t)T * W + R];
                    double BL = pl[(size_t)Bm * W + L], BR = pl[(size_t)Bm * W + R];
                    bx += g * (gam1 * (TR - TL) + (1 - gam1) * (BR - BL));   /* :165-173 */
                    by += g * (gam2 * (BL - TL) + (1 - gam2) * (BR - TR));   /* :182-190 */
                }
                gi2[((size_t)b * 2 + 0) * HW + pix] = bx;
                gi2[((size_t)b * 2 + 1) * HW + pix] = by;
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SeparableConv   my_package/SeparableConv/separableconv_cuda_kernel.cu:29-81 (fwd), :85-135 (bwd)
 *   input1 [B,C,H,W]; input2 (vertical) and input3 (horizontal) [B,F,Ho,Wo]; output [B,C,Ho,Wo]
 *   Ho = H-F+1, Wo = W-F+1.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API int oracle_sepconv_forward(const float *input1, const float *input2, const float *input3,
                                      double *output, int B, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    if (Ho <= 0 || Wo <= 0) return 1;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < Ho; ++h_i)
            for (int w_i = 0; w_i < Wo; ++w_i) {
                const size_t po = (size_t)h_i * Wo + w_i;
                for (int c = 0; c < C; ++c) {
                    const float *pl = input1 + ((size_t)b * C + c) * HW;
                    double out = 0.0;
                    for (int y = 0; y < F; ++y)
                        for (int x = 0; x < F; ++x) {
                            double t1 = pl[(size_t)(h_i + y) * W + (w_i + x)];
                            double t2 = input2[((size_t)b * F + y) * HWo + po];
                            double t3 = input3[((size_t)b * F + x) * HWo + po];
                            out += t1 * t2 * t3;   /* :73-77 */
                        }
                    output[((size_t)b * C + c) * HWo + po] = out;
                }
            }
    return 0;
}

ORACLE_API int oracle_sepconv_backward(const float *input1, const float *input2, const float *input3,
                                       const float *gradoutput, double *gi1, double *gi2, double *gi3,
                                       int B, int C, int H, int W, int F)
{
    const int Ho = H - F + 1, Wo = W - F + 1;
    if (Ho <= 0 || Wo <= 0) return 1;
    const size_t HW = (size_t)H * W, HWo = (size_t)Ho * Wo;
    memset(gi1, 0, sizeof(double) * (size_t)B * C * HW);
    memset(gi2, 0, sizeof(double) * (size_t)B * F * HWo);
    memset(gi3, 0, sizeof(double) * (size_t)B * F * HWo);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int h_i = 0; h_i < Ho; ++h_i)
            for (int w_i = 0; w_i < Wo; ++w_i) {
                const size_t po = (size_t)h_i * Wo + w_i;
                for (int c = 0; c < C; ++c) {
                    const float *pl = input1 + ((size_t)b * C + c) * HW;
                    double g = gradoutput[((size_t)b * C + c) * HWo + po];
                    for (int y = 0; y < F; ++y)
                        for (int x = 0; x < F; ++x) {
                            double t1 = pl[(size_t)(h_i + y) * W + (w_i + x)];
                            double t2 = input2[((size_t)b * F + y) * HWo + po];
                            double t3 = input3[((size_t)b * F + x) * HWo + po];
                            gi1[((size_t)b * C + c) * HW + (size_t)(h_i + y) * W + (w_i + x)] += g * t2 * t3; /* :122-123 */
                            gi2[((size_t)b * F + y) * HWo + po] += g * t1 * t3;                                  /* :124-125 */
                            gi3[((size_t)b * F + x) * HWo + po] += g * t1 * t2;                                  /* :126-127 */
                        }
                }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SeparableConvFlow  my_package/SeparableConvFlow/separableconvflow_cuda_kernel.cu:29-93 (fwd), :97-174 (bwd)
 *   input2, input3 [B,F,Ho,Wo]; flow_output [B,2,Ho,Wo] (channel 0 = x from input3, 1 = y from input2)
 *   The sums and the |sum| > 0 predicate are float32 (they select the -2000 sentinel); the value is float64.
 * ------------------------------------------------------------------------------------------ */
ORACLE_API int oracle_sepconvflow_forward(const float *input2, const float *input3, double *flow,
                                          int B, int Ho, int Wo, int F)
{
    const size_t HWo = (size_t)Ho * Wo;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t po = 0; po < HWo; ++po) {
            const float *src[2] = { input3, input2 };   /* channel 0 <- input3 (x), channel 1 <- input2 (y) */
            for (int ch = 0; ch < 2; ++ch) {
                double num = 0.0, den = 0.0; volatile float denf = 0.0f;
                for (int k = 0; k < F; ++k) {
                    float t = src[ch][((size_t)b * F + k) * HWo + po];
                    num += (double)k * t; den += t; denf = denf + t;   /* :60-64 */
                }
                double v = num / den - ((double)F - 1.0) / 2.0;        /* :66 */
                flow[((size_t)b * 2 + ch) * HWo + po] = (fabsf(denf) > 0.0f) ? v : -2000.0;  /* :68-69 */
            }
        }
    return 0;
}

ORACLE_API int oracle_sepconvflow_backward(const float *input2, const float *input3, const float *gradflow,
                                           double *gi2, double *gi3, int B, int Ho, int Wo, int F)
{
    const size_t HWo = (size_t)Ho * Wo;
    memset(gi2, 0, sizeof(double) * (size_t)B * F * HWo);
    memset(gi3, 0, sizeof(double) * (size_t)B * F * HWo);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (size_t po = 0; po < HWo; ++po) {
            const float *src[2] = { input3, input2 };
            double *dst[2] = { gi3, gi2 };
            for (int ch = 0; ch < 2; ++ch) {
                double num = 0.0, den = 0.0; volatile float denf = 0.0f;
                for (int k = 0; k < F; ++k) {
                    float t = src[ch][((size_t)b * F + k) * HWo + po];
                    num += (double)k * t; den += t; denf = denf + t;
                }
                if (fabsf(denf) > 0.0f) {                               /* :135, :158 */
                    double g = gradflow[((size_t)b * 2 + ch) * HWo + po];
                    double offset = num / (den * den);                   /* :138 */
                    for (int k = 0; k < F; ++k)
                        dst[ch][((size_t)b * F + k) * HWo + po] = g * ((double)k / den - offset);  /* :140-143 */
                }
            }
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Correlation  PWCNet/correlation_package_pytorch1_0/correlation_cuda_kernel.cu:47-334,
 *              output-shape rules correlation_cuda.cc:23-36.
 * The reference repacks to zero-padded NHWC (channels_first, :47-70); the restatement reads the
 * NCHW input through a zero-padding accessor, which is the same function.
 * Returns 0 on success (the reference returns 1 on success, correlation_cuda_kernel.cu:417-426).
 * ------------------------------------------------------------------------------------------ */
static inline double corr_padded(const float *in, int C, int H, int W, int pad, int n, int c, int py, int px)
{
    int y = py - pad, x = px - pad;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0;
    return in[(((size_t)n * C + c) * H + y) * W + x];
}

ORACLE_API void oracle_correlation_outshape(int H, int W, int pad, int k, int md, int s1, int s2,
                                            int *oc, int *oh, int *ow)
{
    int kr = (k - 1) / 2, border = kr + md;                       /* correlation_cuda.cc:23-24 */
    int pH = H + 2 * pad, pW = W + 2 * pad;
    int dr = md / s2;
    *oc = (dr * 2 + 1) * (dr * 2 + 1);                             /* :29 */
    *oh = (int)ceilf((float)(pH - 2 * border) / (float)s1);       /* :31 */
    *ow = (int)ceilf((float)(pW - 2 * border) / (float)s1);       /* :32 */
}

ORACLE_API int oracle_correlation_forward(const float *input1, const float *input2, double *output,
                                          int B, int C, int H, int W,
                                          int pad, int k, int md, int s1, int s2)
{
    int oc, oh, ow; oracle_correlation_outshape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
    if (oh <= 0 || ow <= 0) return 1;
    const int kr = (k - 1) / 2, dr = md / s2, ds = 2 * dr + 1;
    const double nelems = (double)k * k * C;                        /* :104 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
        for (int by = 0; by < oh; ++by)
            for (int bx = 0; bx < ow; ++bx) {
                int y1 = by * s1 + md, x1 = bx * s1 + md;          /* :92-93 (padded coordinates) */
                for (int tj = -dr; tj <= dr; ++tj)
                    for (int ti = -dr; ti <= dr; ++ti) {
                        int x2 = x1 + ti * s2, y2 = y1 + tj * s2;   /* :109-110 */
                        double acc = 0.0;
                        for (int j = -kr; j <= kr; ++j)
                            for (int i = -kr; i <= kr; ++i)
                                for (int ch = 0; ch < C; ++ch)
                                    acc += corr_padded(input1, C, H, W, pad, n, ch, y1 + j, x1 + i) *
                                           corr_padded(input2, C, H, W, pad, n, ch, y2 + j, x2 + i);   /* :116-124 */
                        int tc = (tj + dr) * ds + (ti + dr);        /* :138 */
                        output[(((size_t)n * oc + tc) * oh + by) * ow + bx] = acc / nelems;   /* :143 */
                    }
            }
    return 0;
}

/* C integer division truncates toward zero; the reference relies on it (:172-175). */
ORACLE_API int oracle_correlation_backward(const float *input1, const float *input2, const float *gradoutput,
                                           double *gi1, double *gi2,
                                           int B, int C, int H, int W,
                                           int pad, int k, int md, int s1, int s2)
{
    int oc, oh, ow; oracle_correlation_outshape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
    if (oh <= 0 || ow <= 0) return 1;
    const int kr = (k - 1) / 2, dr = md / s2, ds = 2 * dr + 1;
    const double nelems = (double)k * k * C;
    memset(gi1, 0, sizeof(double) * (size_t)B * C * H * W);
    memset(gi2, 0, sizeof(double) * (size_t)B * C * H * W);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c)
            for (int by = 0; by < H; ++by)
                for (int bx = 0; bx < W; ++bx) {
                    int y = by * s1 + pad, x = bx * s1 + pad;       /* :162-163, grid (H,W,C) :519 */
                    /* ---- input1 (:151-241) ---- */
                    {
                        int xmin = (x - kr - md) / s1, ymin = (y - kr - md) / s1;
                        int xmax = (x + kr - md) / s1, ymax = (y + kr - md) / s1;
                        if (!(xmax < 0 || ymax < 0 || xmin >= ow || ymin >= oh) && !(xmin > xmax || ymin > ymax)) {
                            xmin = imax(0, xmin); xmax = imin(ow - 1, xmax);
                            ymin = imax(0, ymin); ymax = imin(oh - 1, ymax);
                            double sum = 0.0;
                            for (int tc = 0; tc < oc; ++tc) {
                                int i2 = (tc % ds - dr) * s2, j2 = (tc / ds - dr) * s2;   /* :209-210 */
                                double val2 = corr_padded(input2, C, H, W, pad, n, c, y + j2, x + i2);
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i)
                                        sum += (double)gradoutput[(((size_t)n * oc + tc) * oh + j) * ow + i] * val2;
                            }
                            /* index (y - pad, x - pad) in the H x W plane (:238) -- only valid for s1 == 1
                             * or small planes; out-of-plane targets are skipped here */
                            int oy = y - pad, ox = x - pad;
                            if (oy >= 0 && oy < H && ox >= 0 && ox < W)
                                gi1[(((size_t)n * C + c) * H + oy) * W + ox] = sum / nelems;
                        }
                    }
                    /* ---- input2 (:244-334) ---- */
                    {
                        double sum = 0.0;
                        for (int tc = 0; tc < oc; ++tc) {
                            int i2 = (tc % ds - dr) * s2, j2 = (tc / ds - dr) * s2;
                            int xmin = (x - kr - md - i2) / s1, ymin = (y - kr - md - j2) / s1;   /* :291-294 */
                            int xmax = (x + kr - md - i2) / s1, ymax = (y + kr - md - j2) / s1;
                            if (xmax < 0 || ymax < 0 || xmin >= ow || ymin >= oh) continue;
                            if (xmin > xmax || ymin > ymax) continue;
                            xmin = imax(0, xmin); xmax = imin(ow - 1, xmax);
                            ymin = imax(0, ymin); ymax = imin(oh - 1, ymax);
                            double val1 = corr_padded(input1, C, H, W, pad, n, c, y - j2, x - i2);   /* :310-311 */
                            for (int j = ymin; j <= ymax; ++j)
                                for (int i = xmin; i <= xmax; ++i)
                                    sum += (double)gradoutput[(((size_t)n * oc + tc) * oh + j) * ow + i] * val1;
                        }
                        int oy = y - pad, ox = x - pad;
                        if (oy >= 0 && oy < H && ox >= 0 && ox < W)
                            gi2[(((size_t)n * C + c) * H + oy) * W + ox] = sum / nelems;
                    }
                }
    return 0;
}

ORACLE_API int oracle_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
