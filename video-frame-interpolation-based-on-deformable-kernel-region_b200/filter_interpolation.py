"""FilterInterpolation -- adaptive warping with learned per-pixel F x F filters, and the three
deformable-kernel-region (DKR) families, behind the reference's autograd.Function / Module API.

Reference surface mirrored here (citations relative to /root/reference/):
  my_package/FilterInterpolation/FilterInterpolationLayer.py:10-91   FilterInterpolationLayer.apply(input1,input2,input3)
  my_package/FilterInterpolation/FilterInterpolationModule.py:8-20   FilterInterpolationModule()(input1,input2,input3[,input4])
The reference selects the kernel family by (un)commenting lines (FilterInterpolationLayer.py:35-38);
here each family is its own Function and FilterInterpolationModule takes `variant=`.
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._common import check_input, ptr, stream_ptr

__all__ = ["FilterInterpolationLayer", "FilterInterpolationLayerDKR", "FilterInterpolationLayerDeforConv",
           "FilterInterpolationLayerNoFilterWithDeforConv", "FilterInterpolationModule"]


def _filter_size(channels: int) -> int:
    # filter_size = (int) sqrt((float) input3.size(1))   (filterinterpolation_cuda.cc:556-557)
    return int(math.sqrt(float(channels)))


def _check_shapes(input1, input2, input3, input4=None, offsets_in_input3=False):
    B, C, H, W = input1.shape
    # the .cc glue returns error 1 for these (filterinterpolation_cuda.cc:545-553); here they raise
    if input2.shape != (B, 2, H, W):
        raise _lib.VfidkrError(f"input2 must be [B,2,H,W] = {(B, 2, H, W)}, got {tuple(input2.shape)}")
    if input3.dim() != 4 or input3.shape[0] != B or input3.shape[2:] != (H, W):
        raise _lib.VfidkrError(f"input3 must be [B,K,H,W] with B,H,W = {(B, H, W)}, got {tuple(input3.shape)}")
    if offsets_in_input3:
        F = int(math.sqrt(float(input3.shape[1] // 2)))   # filterinterpolation_cuda.cc:395
        if input3.shape[1] != 2 * F * F:
            raise _lib.VfidkrError("input3 must hold 2*F*F offset channels")
    else:
        F = _filter_size(input3.shape[1])
        if input3.shape[1] != F * F:
            raise _lib.VfidkrError("input3 must hold F*F filter channels")
    if input4 is not None and input4.shape != (B, 2 * F * F, H, W):
        raise _lib.VfidkrError(f"input4 must be [B,2*F*F,H,W] = {(B, 2 * F * F, H, W)}, got {tuple(input4.shape)}")
    return B, C, H, W, F


class FilterInterpolationLayer(Function):
    """Live "_ori" family: FilterInterpolationLayer_gpu_forward_ori / _backward_ori."""

    @staticmethod
    def forward(ctx, input1, input2, input3):
        for t, n in ((input1, "input1"), (input2, "input2"), (input3, "input3")):
            check_input(t, n)
        B, C, H, W, F = _check_shapes(input1, input2, input3)
        output = torch.empty_like(input1)   # every element is written by the kernel; no zero-fill needed
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_filterinterpolation_forward_ori", ptr(input1), ptr(input2), ptr(input3), ptr(output),
                      B, C, H, W, F, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, input3)
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, input3 = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, C, H, W = input1.shape
        F = _filter_size(input3.shape[1])
        gi1, gi2, gi3 = torch.empty_like(input1), torch.empty_like(input2), torch.empty_like(input3)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(input1), ptr(input2), ptr(input3),
                      ptr(gradoutput), ptr(gi1), ptr(gi2), ptr(gi3), B, C, H, W, F, stream_ptr(input1.device))
        return gi1, gi2, gi3


def _make_four_input(name: str, suffix: str, doc: str):
    class _Layer(Function):
        @staticmethod
        def forward(ctx, input1, input2, input3, input4):
            for t, n in ((input1, "input1"), (input2, "input2"), (input3, "input3"), (input4, "input4")):
                check_input(t, n)
            B, C, H, W, F = _check_shapes(input1, input2, input3, input4)
            output = torch.empty_like(input1)
            with torch.cuda.device(input1.device):
                _lib.call(f"vfidkr_filterinterpolation_forward_{suffix}", ptr(input1), ptr(input2), ptr(input3),
                          ptr(input4), ptr(output), B, C, H, W, F, stream_ptr(input1.device))
            ctx.save_for_backward(input1, input2, input3, input4)
            return output

        @staticmethod
        def backward(ctx, gradoutput):
            input1, input2, input3, input4 = ctx.saved_tensors
            gradoutput = gradoutput.contiguous()
            B, C, H, W = input1.shape
            F = _filter_size(input3.shape[1])
            gi1, gi2 = torch.empty_like(input1), torch.empty_like(input2)
            gi3, gi4 = torch.empty_like(input3), torch.empty_like(input4)
            with torch.cuda.device(input1.device):
                _lib.call(f"vfidkr_filterinterpolation_backward_{suffix}", ptr(input1), ptr(input2), ptr(input3),
                          ptr(input4), ptr(gradoutput), ptr(gi1), ptr(gi2), ptr(gi3), ptr(gi4),
                          B, C, H, W, F, stream_ptr(input1.device))
            return gi1, gi2, gi3, gi4

    _Layer.__name__ = _Layer.__qualname__ = name
    _Layer.__doc__ = doc
    return _Layer


FilterInterpolationLayerDKR = _make_four_input(
    "FilterInterpolationLayerDKR", "dkr",
    "4-input DKR family with static quadrants: FilterInterpolationLayer_gpu_forward / _gpu_backward "
    "(FilterInterpolationLayer.py:36,73; filterinterpolation_cuda.cc:11-187). input4 = [B,2*F*F,H,W] offsets.")
FilterInterpolationLayerDeforConv = _make_four_input(
    "FilterInterpolationLayerDeforConv", "deforconv",
    "DKR with data-dependent quadrants: ..._gpu_forward_deforconv / _gpu_backward_deforconv "
    "(FilterInterpolationLayer.py:37,74; filterinterpolation_cuda.cc:191-367).")


class FilterInterpolationLayerNoFilterWithDeforConv(Function):
    """DKR with unit filter weights; input3 is the [B,2*F*F,H,W] offset field
    (FilterInterpolationLayer.py:38,75; filterinterpolation_cuda.cc:374-533)."""

    @staticmethod
    def forward(ctx, input1, input2, input3):
        for t, n in ((input1, "input1"), (input2, "input2"), (input3, "input3")):
            check_input(t, n)
        B, C, H, W, F = _check_shapes(input1, input2, input3, offsets_in_input3=True)
        output = torch.empty_like(input1)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_filterinterpolation_forward_nofilterwithdeforconv", ptr(input1), ptr(input2),
                      ptr(input3), ptr(output), B, C, H, W, F, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, input3)
        ctx.filter_size = F
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, input3 = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, C, H, W = input1.shape
        gi1, gi2, gi3 = torch.empty_like(input1), torch.empty_like(input2), torch.empty_like(input3)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_filterinterpolation_backward_nofilterwithdeforconv", ptr(input1), ptr(input2),
                      ptr(input3), ptr(gradoutput), ptr(gi1), ptr(gi2), ptr(gi3), B, C, H, W, ctx.filter_size,
                      stream_ptr(input1.device))
        return gi1, gi2, gi3


_THREE_INPUT = {"ori": FilterInterpolationLayer, "nofilterwithdeforconv": FilterInterpolationLayerNoFilterWithDeforConv}
_FOUR_INPUT = {"dkr": FilterInterpolationLayerDKR, "deforconv": FilterInterpolationLayerDeforConv}


class FilterInterpolationBlendLayer(Function):
    """Warps both temporal directions with the "_ori" family and blends them in the same two launches:
    w0 * FI(ref0, offset0, filter0) + w2 * FI(ref2, offset2, filter2) -- `ref0/2 + ref2/2` of networks/DAIN.py:573
    (w0 = w2 = 0.5) and `(1-t)*ref0 + t*ref2` of DAIN_slowmotion.py:335 -- without the two intermediate frames and the
    blend pass.  The backward runs the two "_ori" backward kernels on the scaled gradient."""

    @staticmethod
    def forward(ctx, ref0, ref2, offset0, offset2, filter0, filter2, w0=0.5, w2=0.5, out=None):
        for t, n in ((ref0, "ref0"), (ref2, "ref2"), (offset0, "offset0"), (offset2, "offset2"),
                     (filter0, "filter0"), (filter2, "filter2")):
            check_input(t, n)
        B, C, H, W, F = _check_shapes(ref0, offset0, filter0)
        if _check_shapes(ref2, offset2, filter2) != (B, C, H, W, F):
            raise _lib.VfidkrError("both directions must have the same shapes")
        if out is None:
            output, bs = torch.empty_like(ref0), 0
        else:
            # a channel slice of a wider contiguous tensor (e.g. the rectify-input concat): channels H*W apart, any batch stride
            if out.shape != ref0.shape or out.dtype != torch.float32 or out.device != ref0.device or \
                    tuple(out.stride()[1:]) != (H * W, W, 1) or out.stride(0) < C * H * W:
                raise _lib.VfidkrError("out must be a float32 [B,C,H,W] view with strides (>= C*H*W, H*W, W, 1) on the inputs' device")
            output, bs = out, out.stride(0)
        with torch.cuda.device(ref0.device):
            sp = stream_ptr(ref0.device)
            _lib.call("vfidkr_filterinterpolation_forward_ori_blend", ptr(ref0), ptr(offset0), ptr(filter0), output.data_ptr(),
                      B, C, H, W, F, float(w0), 0, bs, sp)
            _lib.call("vfidkr_filterinterpolation_forward_ori_blend", ptr(ref2), ptr(offset2), ptr(filter2), output.data_ptr(),
                      B, C, H, W, F, float(w2), 1, bs, sp)
        if out is not None:
            ctx.mark_dirty(out)
        ctx.save_for_backward(ref0, ref2, offset0, offset2, filter0, filter2)
        ctx.weights = (float(w0), float(w2))
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        ref0, ref2, offset0, offset2, filter0, filter2 = ctx.saved_tensors
        B, C, H, W = ref0.shape
        F = _filter_size(filter0.shape[1])
        grads = []
        with torch.cuda.device(ref0.device):
            for img, off, flt, wgt in ((ref0, offset0, filter0, ctx.weights[0]), (ref2, offset2, filter2, ctx.weights[1])):
                g = (gradoutput * wgt).contiguous()
                gi1, gi2, gi3 = torch.empty_like(img), torch.empty_like(off), torch.empty_like(flt)
                _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(img), ptr(off), ptr(flt), ptr(g),
                          ptr(gi1), ptr(gi2), ptr(gi3), B, C, H, W, F, stream_ptr(img.device))
                grads.append((gi1, gi2, gi3))
        (a1, a2, a3), (b1, b2, b3) = grads
        return a1, b1, a2, b2, a3, b3, None, None, None


def filter_interpolate_blend(ref0, ref2, offset0, offset2, filter0, filter2, w0=0.5, w2=0.5, out=None):
    """w0 * FilterInterpolation(ref0, offset0, filter0) + w2 * FilterInterpolation(ref2, offset2, filter2), fused.
    `out` (inference): a [B,C,H,W] channel slice of a wider contiguous tensor to write into, e.g. `cat[:, 3:6]`."""
    if out is not None and torch.is_grad_enabled() and any(t.requires_grad for t in (ref0, ref2, offset0, offset2, filter0, filter2)):
        raise _lib.VfidkrError("out= is for inference (no_grad): writing into a slice of another tensor is not differentiable here")
    return FilterInterpolationBlendLayer.apply(ref0, ref2, offset0, offset2, filter0, filter2, w0, w2, out)


def filter_interpolate_into(input1, input2, input3, out, scale=1.0, accumulate=False):
    """Inference helper: out = scale * FilterInterpolation(input1, input2, input3) (+ out), written straight into `out`,
    a [B,C,H,W] channel slice of a wider contiguous tensor -- e.g. the warped context features ctx0 / ctx2 (C = 196) into
    their slices of the 437-channel rectify input (DAIN_slowmotion.py:167-181) without the intermediate tensors and the
    torch.cat.  Any channel count (the many-channel kernel carries the same epilogue as the C <= 4 ones)."""
    for t, n in ((input1, "input1"), (input2, "input2"), (input3, "input3")):
        check_input(t, n)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (input1, input2, input3)):
        raise _lib.VfidkrError("filter_interpolate_into is for inference (no_grad): writing into a slice is not differentiable here")
    B, C, H, W, F = _check_shapes(input1, input2, input3)
    if out.shape != input1.shape or out.dtype != torch.float32 or out.device != input1.device or \
            tuple(out.stride()[1:]) != (H * W, W, 1) or out.stride(0) < C * H * W:
        raise _lib.VfidkrError("out must be a float32 [B,C,H,W] view with strides (>= C*H*W, H*W, W, 1) on the inputs' device")
    with torch.cuda.device(input1.device):
        _lib.call("vfidkr_filterinterpolation_forward_ori_blend", ptr(input1), ptr(input2), ptr(input3), out.data_ptr(),
                  B, C, H, W, F, float(scale), 1 if accumulate else 0, out.stride(0), stream_ptr(input1.device))
    return out


class FilterInterpolationModule(Module):
    """FilterInterpolationModule()(input1, input2, input3)            -> "_ori" (FilterInterpolationModule.py:13-17)
    FilterInterpolationModule()(input1, input2, input3, input4)      -> 4-input DKR (:18-20)
    `variant` picks the family explicitly: "ori" | "nofilterwithdeforconv" (3 inputs), "dkr" | "deforconv" (4 inputs)."""

    def __init__(self, variant: str | None = None):
        super().__init__()
        if variant is not None and variant not in _THREE_INPUT and variant not in _FOUR_INPUT:
            raise ValueError(f"unknown FilterInterpolation variant {variant!r}")
        self.variant = variant

    def forward(self, input1, input2, input3, input4=None):
        if input4 is None:
            layer = _THREE_INPUT[self.variant or "ori"]
            return layer.apply(input1, input2, input3)
        layer = _FOUR_INPUT[self.variant or "dkr"]
        return layer.apply(input1, input2, input3, input4)
