// fi_common.cuh -- device helpers shared by the FilterInterpolation kernels (filterinterpolation.cu, fi_strip.cu):
// the float32 index arithmetic of the reference, written once so every kernel truncates and compares identically.
#pragma once

#include "common.cuh"

namespace vfidkr {

enum { V_ORI = 0, V_DKR = 1, V_DEFOR = 2, V_NOFILT = 3 };

struct FiPix {
    bool in_range;
    int ix, iy, L, T;
    float x2, y2, alpha, beta;
};

// range test + window origin, float32 exactly as the reference writes it (:2731-2743)
__device__ __forceinline__ FiPix fi_pixel(int w_i, int h_i, float fx, float fy, int W, int H, int F)
{
    FiPix p;
    p.x2 = __fadd_rn((float)w_i, fx);
    p.y2 = __fadd_rn((float)h_i, fy);
    p.in_range = p.x2 >= 0.0f && p.y2 >= 0.0f && p.x2 <= (float)(W - 1) && p.y2 <= (float)(H - 1) &&
                 fabsf(fx) < (float)W / 2.0f && fabsf(fy) < (float)H / 2.0f;
    p.ix = (int)p.x2;
    p.iy = (int)p.y2;
    p.L = p.ix + 1 - F / 2;
    p.T = p.iy + 1 - F / 2;
    p.alpha = __fsub_rn(p.x2, (float)p.ix);
    p.beta = __fsub_rn(p.y2, (float)p.iy);
    return p;
}

// one deformed tap: the four read offsets inside a channel plane and the bilinear fractions
struct Deform {
    int aTL, aTR, aBL, aBR;
    float phiX, phiY;
    bool top, left;  // data-dependent quadrant (fracY <= y2, fracX <= x2)
};

__device__ __forceinline__ Deform fi_deform(int cy, int cx, float offY, float offX, const FiPix &p, int H, int W)
{
    Deform d;
    const float fracY = __fadd_rn((float)cy, offY);  // :98
    const float fracX = __fadd_rn((float)cx, offX);  // :99
    const int Top = (int)fracY, Left = (int)fracX;   // :102-103 (cvt.rzi saturates, NaN -> 0)
    d.phiY = __fsub_rn(fracY, (float)Top);           // :100
    d.phiX = __fsub_rn(fracX, (float)Left);          // :101
    // The reference leaves Top/Left/Bottom/Right unclamped (undefined behaviour outside the plane);
    // reads are clamped to the plane here, weights are untouched ("in-contract domain").
    const int t = clampi(Top, 0, H - 1), b = clampi(min(Top, H - 1) + 1, 0, H - 1);
    const int l = clampi(Left, 0, W - 1), r = clampi(min(Left, W - 1) + 1, 0, W - 1);
    d.aTL = t * W + l; d.aTR = t * W + r; d.aBL = b * W + l; d.aBR = b * W + r;
    d.top = fracY <= p.y2;
    d.left = fracX <= p.x2;
    return d;
}

// per-tap blend coefficients selected by the tap's quadrant (0=TL 1=TR 2=BL 3=BR):
//   q  : weight of the quadrant sum in the output            (:2789-2793)
//   cx : coefficient of the tap in d(out)/d(flow_x)  = gamma*(TR-TL) + (1-gamma)*(BR-BL), gamma = 1-beta   (:2965-3013)
//   cy : coefficient of the tap in d(out)/d(flow_y)  = gamma'*(BL-TL) + (1-gamma')*(BR-TR), gamma' = 1-alpha (:3036-3084)
struct QuadCoef { float q, cx, cy; };
__device__ __forceinline__ QuadCoef quad_coef(bool top, bool left, float alpha, float beta)
{
    QuadCoef r;
    const float ax = left ? 1.0f - alpha : alpha;   // x-factor of q
    const float by = top ? 1.0f - beta : beta;      // y-factor of q
    r.q = top ? (left ? (1 - alpha) * (1 - beta) : alpha * (1 - beta)) : (left ? (1 - alpha) * beta : alpha * beta);
    const float gam = 1.0f - beta, gam2 = 1.0f - alpha;
    r.cx = (left ? -1.0f : 1.0f) * (top ? gam : 1.0f - gam);
    r.cy = (top ? -1.0f : 1.0f) * (left ? gam2 : 1.0f - gam2);
    (void)ax; (void)by;
    return r;
}

}  // namespace vfidkr
