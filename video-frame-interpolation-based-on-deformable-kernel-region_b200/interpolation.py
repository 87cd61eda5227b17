"""Interpolation / InterpolationCh -- plain bilinear backward warp, zero fill out of range.

Reference surface (relative to /root/reference/):
  my_package/Interpolation/InterpolationLayer.py:10-77, InterpolationModule.py   InterpolationModule()(input1, input2)
  my_package/InterpolationCh/InterpolationChLayer.py, InterpolationChModule.py   InterpolationChModule(ch)(input1, input2)
Interpolation enforces C == 3 (interpolation_cuda.cc:19); InterpolationCh does not.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._common import check_input, ptr, stream_ptr

__all__ = ["InterpolationLayer", "InterpolationModule", "InterpolationChLayer", "InterpolationChModule"]


def _forward(ctx, input1, input2, require_c3):
    check_input(input1, "input1")
    check_input(input2, "input2")
    B, C, H, W = input1.shape
    if input2.shape != (B, 2, H, W):   # interpolation_cuda.cc:21-27
        raise _lib.VfidkrError(f"input2 must be [B,2,H,W] = {(B, 2, H, W)}, got {tuple(input2.shape)}")
    output = torch.empty_like(input1)
    with torch.cuda.device(input1.device):
        _lib.call("vfidkr_interpolation_forward", ptr(input1), ptr(input2), ptr(output), B, C, H, W, require_c3,
                  stream_ptr(input1.device))
    ctx.save_for_backward(input1, input2)
    return output


def _backward(ctx, gradoutput, require_c3):
    input1, input2 = ctx.saved_tensors
    gradoutput = gradoutput.contiguous()
    B, C, H, W = input1.shape
    gradinput1, gradinput2 = torch.empty_like(input1), torch.empty_like(input2)
    with torch.cuda.device(input1.device):
        _lib.call("vfidkr_interpolation_backward", ptr(input1), ptr(input2), ptr(gradoutput), ptr(gradinput1),
                  ptr(gradinput2), B, C, H, W, require_c3, stream_ptr(input1.device))
    return gradinput1, gradinput2


class InterpolationLayer(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        return _forward(ctx, input1, input2, 1)

    @staticmethod
    def backward(ctx, gradoutput):
        return _backward(ctx, gradoutput, 1)


class InterpolationChLayer(Function):
    @staticmethod
    def forward(ctx, input1, input2):
        return _forward(ctx, input1, input2, 0)

    @staticmethod
    def backward(ctx, gradoutput):
        return _backward(ctx, gradoutput, 0)


class InterpolationModule(Module):
    def forward(self, input1, input2):
        return InterpolationLayer.apply(input1, input2)


class InterpolationChModule(Module):
    def __init__(self, ch):
        super().__init__()
        self.ch = ch

    def forward(self, input1, input2):
        return InterpolationChLayer.apply(input1, input2)
