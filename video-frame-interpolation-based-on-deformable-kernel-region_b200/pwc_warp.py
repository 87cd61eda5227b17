"""PWCDCNet.warp as one native operator.

The reference builds the sampling grid, normalises it, calls torch.nn.functional.grid_sample twice (features and an
all-ones tensor) and thresholds the second result into a validity mask (PWCNet/PWCNet.py:159-199); it also caps
B <= 3, H <= 1024, W <= 2048 through a pre-allocated grid (:142-155).  `pwc_warp(x, flo)` computes the same values in
one kernel per direction of the autograd graph, with no caps.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib
from ._common import check_input, ptr, stream_ptr


class PWCWarpLayer(Function):
    @staticmethod
    def forward(ctx, x, flo):
        check_input(x, "x")
        check_input(flo, "flo")
        B, C, H, W = x.shape
        if flo.shape != (B, 2, H, W):
            raise _lib.VfidkrError(f"flo must be [B,2,H,W] = {(B, 2, H, W)}, got {tuple(flo.shape)}")
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call("vfidkr_pwcwarp_forward", ptr(x), ptr(flo), ptr(out), B, C, H, W, stream_ptr(x.device))
        ctx.save_for_backward(x, flo)
        return out

    @staticmethod
    def backward(ctx, gradoutput):
        x, flo = ctx.saved_tensors
        B, C, H, W = x.shape
        gradoutput = gradoutput.contiguous()
        gx, gf = torch.empty_like(x), torch.empty_like(flo)
        with torch.cuda.device(x.device):
            _lib.call("vfidkr_pwcwarp_backward", ptr(x), ptr(flo), ptr(gradoutput), ptr(gx), ptr(gf), B, C, H, W,
                      stream_ptr(x.device))
        return gx, gf


def pwc_warp(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    """Drop-in for PWCDCNet.warp(x, flo): warp x [B,C,H,W] back by the flow flo [B,2,H,W], masked."""
    return PWCWarpLayer.apply(x, flo)
