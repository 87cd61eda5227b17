"""SeparableConv / SeparableConvFlow.

Reference surface (relative to /root/reference/):
  my_package/SeparableConv/SeparableConvLayer.py:10-88, SeparableConvModule.py       SeparableConvModule(filtersize)(input1,input2,input3)
  my_package/SeparableConvFlow/SeparableConvFlowLayer.py:10-94, ...Module.py          SeparableConvFlowModule(filtersize)(input1,input2,input3)
The reference layers are legacy instance-style Functions (and SeparableConvLayer imports a module that no
longer exists, :4); they are restated as static Functions with the same call signatures and shape asserts.
"""
from __future__ import annotations

import warnings

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._common import check_input, ptr, stream_ptr

__all__ = ["SeparableConvLayer", "SeparableConvModule", "SeparableConvFlowLayer", "SeparableConvFlowModule"]


def _shapes(input1, input2, input3, filtersize):
    for t, n in ((input1, "input1"), (input2, "input2"), (input3, "input3")):
        check_input(t, n)
    B, C, H, W = input1.shape
    F = min(input2.size(1), input3.size(1))
    Ho = min(input2.size(2), input3.size(2))
    Wo = min(input2.size(3), input3.size(3))
    # SeparableConvLayer.py:24-26
    assert H - filtersize == Ho - 1
    assert W - filtersize == Wo - 1
    assert F == filtersize
    if input2.shape != (B, F, Ho, Wo) or input3.shape != (B, F, Ho, Wo):   # separableconv_cuda.cc:23-31
        raise _lib.VfidkrError("input2/input3 must both be [B,F,H-F+1,W-F+1]")
    return B, C, H, W, F, Ho, Wo


class SeparableConvLayer(Function):
    @staticmethod
    def forward(ctx, input1, input2, input3, filtersize):
        B, C, H, W, F, Ho, Wo = _shapes(input1, input2, input3, filtersize)
        if C != 3:   # separableconv_cuda.cc:21
            raise _lib.VfidkrError("SeparableConv requires 3 input channels")
        output = torch.empty((B, C, Ho, Wo), dtype=input1.dtype, device=input1.device)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_separableconv_forward", ptr(input1), ptr(input2), ptr(input3), ptr(output),
                      B, C, H, W, F, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, input3)
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, input3 = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, C, H, W = input1.shape
        F = input2.size(1)
        gi1, gi2, gi3 = torch.empty_like(input1), torch.empty_like(input2), torch.empty_like(input3)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_separableconv_backward", ptr(input1), ptr(input2), ptr(input3), ptr(gradoutput),
                      ptr(gi1), ptr(gi2), ptr(gi3), B, C, H, W, F, stream_ptr(input1.device))
        return gi1, gi2, gi3, None


class SeparableConvModule(Module):
    def __init__(self, filtersize):
        super().__init__()
        self.filtersize = filtersize

    def forward(self, input1, input2, input3):
        return SeparableConvLayer.apply(input1, input2, input3, self.filtersize)


class SeparableConvFlowLayer(Function):
    @staticmethod
    def forward(ctx, input1, input2, input3, filtersize):
        B, C, H, W, F, Ho, Wo = _shapes(input1, input2, input3, filtersize)
        flow_output = torch.empty((B, 2, Ho, Wo), dtype=input1.dtype, device=input1.device)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_separableconvflow_forward", ptr(input2), ptr(input3), ptr(flow_output),
                      B, Ho, Wo, F, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, input3)
        return flow_output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, input3 = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, F, Ho, Wo = input2.shape
        # the flow back-propagates nothing to input1 (SeparableConvFlowLayer.py:69)
        gi1 = torch.zeros_like(input1)
        gi2, gi3 = torch.empty_like(input2), torch.empty_like(input3)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_separableconvflow_backward", ptr(input2), ptr(input3), ptr(gradoutput), ptr(gi2),
                      ptr(gi3), B, Ho, Wo, F, stream_ptr(input1.device))
        return gi1, gi2, gi3, None


class SeparableConvFlowModule(Module):
    def __init__(self, filtersize):
        super().__init__()
        self.filtersize = filtersize
        # SeparableConvFlowLayer.py:13
        warnings.warn("\nSeparable Conv Flow Layer is not precise enough for optical flow due to a divison operation")

    def forward(self, input1, input2, input3):
        return SeparableConvFlowLayer.apply(input1, input2, input3, self.filtersize)
