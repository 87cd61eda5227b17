"""Correlation -- FlowNet / PWC-Net cost volume.

Reference surface (relative to /root/reference/):
  PWCNet/correlation_package_pytorch1_0/correlation.py:6-63
      Correlation(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)(input1, input2)
      CorrelationFunction(pad_size, ...)(input1, input2)      (legacy instance-style Function)
The legacy Function is restated as a static Function; `CorrelationFunction(...)` still returns a callable
with the same (input1, input2) signature.  No rbot1/rbot2 scratch tensors are allocated.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._common import check_input, ptr, stream_ptr

__all__ = ["Correlation", "CorrelationFunction", "correlation_output_shape", "correlation_pair"]


def correlation_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    """(channels, height, width) per correlation_cuda.cc:23-36."""
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.call("vfidkr_correlation_outshape", H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
              ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
    return oc.value, oh.value, ow.value


class _CorrelationOp(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        check_input(input1, "input1")
        check_input(input2, "input2")
        if input1.shape != input2.shape or input1.dim() != 4:
            raise _lib.VfidkrError("input1 and input2 must be [B,C,H,W] tensors of the same shape")
        B, C, H, W = input1.shape
        oc, oh, ow = correlation_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
        if oh <= 0 or ow <= 0:
            raise _lib.VfidkrError("correlation output would be empty")
        output = torch.empty((B, oc, oh, ow), dtype=input1.dtype, device=input1.device)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_correlation_forward", ptr(input1), ptr(input2), ptr(output), B, C, H, W,
                      pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                      stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2)
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        B, C, H, W = input1.shape
        g1, g2 = torch.empty_like(input1), torch.empty_like(input2)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_correlation_backward", ptr(input1), ptr(input2), ptr(grad_output), ptr(g1), ptr(g2),
                      B, C, H, W, *ctx.params, stream_ptr(input1.device))
        return g1, g2, None, None, None, None, None, None


class _CorrelationPairOp(Function):
    """Both temporal directions of one pyramid level in ONE launch: (corr(input1, input2), corr(input2, input1)), the
    results of two calls of the single op (vfidkr_correlation_forward_pair; no reference counterpart)."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        check_input(input1, "input1")
        check_input(input2, "input2")
        if input1.shape != input2.shape or input1.dim() != 4:
            raise _lib.VfidkrError("input1 and input2 must be [B,C,H,W] tensors of the same shape")
        B, C, H, W = input1.shape
        oc, oh, ow = correlation_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
        if oh <= 0 or ow <= 0:
            raise _lib.VfidkrError("correlation output would be empty")
        out12 = torch.empty((B, oc, oh, ow), dtype=input1.dtype, device=input1.device)
        out21 = torch.empty_like(out12)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_correlation_forward_pair", ptr(input1), ptr(input2), ptr(out12), ptr(out21), B, C, H, W,
                      pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2)
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        return out12, out21

    @staticmethod
    def backward(ctx, g12, g21):
        input1, input2 = ctx.saved_tensors
        B, C, H, W = input1.shape
        a1, a2, b2, b1 = (torch.empty_like(input1) for _ in range(4))
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_correlation_backward", ptr(input1), ptr(input2), ptr(g12.contiguous()), ptr(a1), ptr(a2),
                      B, C, H, W, *ctx.params, stream_ptr(input1.device))
            _lib.call("vfidkr_correlation_backward", ptr(input2), ptr(input1), ptr(g21.contiguous()), ptr(b2), ptr(b1),
                      B, C, H, W, *ctx.params, stream_ptr(input1.device))
        return a1 + b1, a2 + b2, None, None, None, None, None, None


def correlation_pair(input1, input2, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1):
    """(Correlation(...)(input1, input2), Correlation(...)(input2, input1)) from one launch."""
    return _CorrelationPairOp.apply(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)


class CorrelationFunction:
    """Callable with the reference's constructor defaults (correlation.py:8)."""

    def __init__(self, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        self.pad_size, self.kernel_size, self.max_displacement = pad_size, kernel_size, max_displacement
        self.stride1, self.stride2, self.corr_multiply = stride1, stride2, corr_multiply

    def __call__(self, input1, input2):
        return _CorrelationOp.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                    self.stride1, self.stride2, self.corr_multiply)


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size, self.kernel_size, self.max_displacement = pad_size, kernel_size, max_displacement
        self.stride1, self.stride2, self.corr_multiply = stride1, stride2, corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction(self.pad_size, self.kernel_size, self.max_displacement, self.stride1,
                                   self.stride2, self.corr_multiply)(input1, input2)

    def both_directions(self, input1, input2):
        """(self(input1, input2), self(input2, input1)) from ONE launch (not in the reference's Module)."""
        return correlation_pair(input1, input2, self.pad_size, self.kernel_size, self.max_displacement, self.stride1,
                                self.stride2, self.corr_multiply)
