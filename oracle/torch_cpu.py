"""Vectorised PyTorch-CPU fp32 restatement of the bench step's operators.

TEST / BASELINE INFRASTRUCTURE ONLY, like the rest of oracle/: bench.py's `cpu_baseline` leg times it on the host
cores next to the float64 C oracle, tests/test_oracle_kat.py holds it to that oracle.  It is what a user WITHOUT the
CUDA extensions would write in plain PyTorch -- gather / scatter_add_ / shifted products -- and is never imported by
the product package (which has no CPU path at all).

Reference formulas: FilterInterpolation "_ori" forward filterinterpolation_cuda_kernel.cu:2692-2823; (Depth)FlowProjection
forward incl. averaging and hole filling depthflowprojection_cuda_kernel.cu:29-241; correlation forward
correlation_cuda_kernel.cu:74-147 (pad = md = 4, kernel 1, strides 1).
"""
from __future__ import annotations

import torch


def fi_ori_forward(input1: torch.Tensor, input2: torch.Tensor, input3: torch.Tensor) -> torch.Tensor:
    B, C, H, W = input1.shape
    F = int(round(input3.shape[1] ** 0.5))
    fx, fy = input2[:, 0], input2[:, 1]
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    x2, y2 = xs + fx, ys + fy
    ok = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1) & (fx.abs() < W / 2.0) & (fy.abs() < H / 2.0)     # :2735-2736
    ix, iy = x2.to(torch.int64), y2.to(torch.int64)            # truncation, as (int)
    alpha, beta = x2 - ix, y2 - iy                              # :2742-2743
    L, T = ix + 1 - F // 2, iy + 1 - F // 2
    flat = input1.reshape(B, C, H * W)
    quad = [torch.zeros(B, C, H, W) for _ in range(4)]          # TL, TR, BL, BR
    for fj in range(F):
        yy = (T + fj).clamp(0, H - 1)
        for fi in range(F):
            xx = (L + fi).clamp(0, W - 1)
            idx = (yy * W + xx).view(B, 1, H * W).expand(B, C, H * W)
            v = torch.gather(flat, 2, idx).view(B, C, H, W) * input3[:, fj * F + fi].unsqueeze(1)
            quad[(2 if fj >= F // 2 else 0) + (1 if fi >= F // 2 else 0)] += v     # split at int(x2), int(y2) (:2749-2787)
    a, b = alpha.unsqueeze(1), beta.unsqueeze(1)
    out = (1 - a) * (1 - b) * quad[0] + a * (1 - b) * quad[1] + (1 - a) * b * quad[2] + a * b * quad[3]
    return torch.where(ok.unsqueeze(1), out, input1)            # out of range: copy (:2814-2819)


def _nearest_valid(valid: torch.Tensor, dim: int, reverse: bool) -> torch.Tensor:
    """Index of the nearest True strictly before (or, reversed, after) each position along `dim`; -1 if none."""
    n = valid.shape[dim]
    shape = [1] * valid.dim()
    shape[dim] = n
    pos = torch.arange(n).view(shape).expand_as(valid)
    if reverse:
        r = _nearest_valid(valid.flip(dim), dim, False)          # in flipped coordinates
        return torch.where(r >= 0, n - 1 - r, r).flip(dim)
    cand = torch.where(valid, pos, torch.full_like(pos, -1))
    best = torch.cummax(cand, dim).values
    return torch.cat([torch.full_like(best.narrow(dim, 0, 1), -1), best.narrow(dim, 0, n - 1)], dim)       # strictly before


def flowprojection_forward(flow: torch.Tensor, depth: torch.Tensor | None, fillhole: int):
    B, _, H, W = flow.shape
    fx, fy = flow[:, 0], flow[:, 1]
    d = depth[:, 0] if depth is not None else torch.ones_like(fx)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    x2, y2 = xs + fx, ys + fy
    ok = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1)
    L, T = x2.to(torch.int64).clamp(0, W - 1), y2.to(torch.int64).clamp(0, H - 1)
    R, Bm = (L + 1).clamp(max=W - 1), (T + 1).clamp(max=H - 1)
    w = torch.where(ok, d, torch.zeros_like(d))
    planes = torch.stack([-w * fx, -w * fy, w], 1).reshape(B, 3, H * W)
    acc = torch.zeros(B, 3, H * W)
    for yy, xx in ((T, L), (T, R), (Bm, L), (Bm, R)):           # a clamped corner is hit twice, as in the reference
        acc.scatter_add_(2, (yy * W + xx).view(B, 1, H * W).expand(B, 3, H * W), planes)
    acc = acc.view(B, 3, H, W)
    count = acc[:, 2:3]
    out = torch.where(count > 0, acc[:, :2] / count.clamp_min(1e-30), acc[:, :2])
    if fillhole:
        valid = count[:, 0] != 0                                 # sources (:175-213); holes: count <= 0
        hole = ~(count[:, 0] > 0)
        num = torch.zeros(B, 2, H, W)
        den = torch.zeros(B, 1, H, W)
        bi = torch.arange(B).view(B, 1, 1)
        for dim, rev in ((2, False), (2, True), (1, False), (1, True)):     # left, right, up, down
            idx = _nearest_valid(valid, dim, rev)
            found = idx >= 0
            j = idx.clamp_min(0)
            yy = ys.to(torch.int64).expand(B, H, W) if dim == 2 else j
            xx = j if dim == 2 else xs.to(torch.int64).expand(B, H, W)
            use = found & (count[bi, 0, yy, xx] > 0)
            num += torch.where(use.unsqueeze(1), out[bi.unsqueeze(1), torch.arange(2).view(1, 2, 1, 1), yy.unsqueeze(1), xx.unsqueeze(1)],
                               torch.zeros(()))
            den += use.unsqueeze(1).float()
        fill = hole.unsqueeze(1) & (den > 0)
        out = torch.where(fill, num / den.clamp_min(1), out)
    return out, count


def correlation_forward(f1: torch.Tensor, f2: torch.Tensor, md: int = 4) -> torch.Tensor:
    B, C, H, W = f1.shape
    p = torch.nn.functional.pad(f2, (md, md, md, md))
    outs = []
    for tj in range(-md, md + 1):
        for ti in range(-md, md + 1):
            outs.append((f1 * p[:, :, md + tj:md + tj + H, md + ti:md + ti + W]).mean(1))
    return torch.stack(outs, 1)
