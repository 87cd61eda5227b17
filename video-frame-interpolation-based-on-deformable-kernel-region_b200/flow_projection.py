"""FlowProjection / DepthFlowProjection -- forward-splatted flow projection.

Reference surface (relative to /root/reference/):
  my_package/FlowProjection/FlowProjectionLayer.py:10-87, FlowProjectionModule.py:5-17
      FlowProjectionModule(requires_grad=True)(input1) -> FlowProjectionLayer.apply(input1, requires_grad)
  my_package/DepthFlowProjection/DepthFlowProjectionLayer.py:7-98, DepthFlowProjectionModule.py:7-16
      DepthFlowProjectionModule(requires_grad=True)(input1, input2)
Hole filling runs only when requires_grad is False, i.e. at inference (FlowProjectionLayer.py:23).
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._common import check_input, ptr, stream_ptr

__all__ = ["FlowProjectionLayer", "FlowProjectionModule", "DepthFlowProjectionLayer", "DepthFlowProjectionModule"]


class FlowProjectionLayer(Function):
    @staticmethod
    def forward(ctx, input1, requires_grad):
        check_input(input1, "input1")
        B, ch, H, W = input1.shape
        if ch != 2:   # flowprojection_cuda.cc:20
            raise _lib.VfidkrError("input1 must be a [B,2,H,W] flow")
        fillhole = 1 if requires_grad == False else 0   # noqa: E712  (FlowProjectionLayer.py:23)
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        output = torch.empty_like(input1)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_flowprojection_forward", ptr(input1), ptr(count), ptr(output), B, H, W, fillhole,
                      stream_ptr(input1.device))
        ctx.save_for_backward(input1, count)
        ctx.fillhole = fillhole
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, count = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, _, H, W = input1.shape
        gradinput1 = torch.empty_like(input1)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_flowprojection_backward", ptr(input1), ptr(count), ptr(gradoutput), ptr(gradinput1),
                      B, H, W, stream_ptr(input1.device))
        return gradinput1, None   # FlowProjectionLayer.py:87


class FlowProjectionModule(Module):
    def __init__(self, requires_grad=True):
        super().__init__()
        self.requires_grad = requires_grad

    def forward(self, input1):
        return FlowProjectionLayer.apply(input1, self.requires_grad)


class DepthFlowProjectionLayer(Function):
    @staticmethod
    def forward(ctx, input1, input2, requires_grad):
        check_input(input1, "input1")
        check_input(input2, "input2")
        B, ch, H, W = input1.shape
        if ch != 2:   # depthflowprojection_cuda.cc:21
            raise _lib.VfidkrError("input1 must be a [B,2,H,W] flow")
        if input2.shape != (B, 1, H, W):   # :28
            raise _lib.VfidkrError(f"input2 must be [B,1,H,W] = {(B, 1, H, W)}, got {tuple(input2.shape)}")
        fillhole = 1 if requires_grad == False else 0   # noqa: E712  (DepthFlowProjectionLayer.py:27)
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        output = torch.empty_like(input1)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_depthflowprojection_forward", ptr(input1), ptr(input2), ptr(count), ptr(output),
                      B, H, W, fillhole, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, count, output)
        ctx.fillhole = fillhole
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, count, output = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, _, H, W = input1.shape
        gradinput1, gradinput2 = torch.empty_like(input1), torch.empty_like(input2)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_depthflowprojection_backward", ptr(input1), ptr(input2), ptr(count), ptr(output),
                      ptr(gradoutput), ptr(gradinput1), ptr(gradinput2), B, H, W, stream_ptr(input1.device))
        return gradinput1, gradinput2, None   # DepthFlowProjectionLayer.py:98


class DepthFlowProjectionModule(Module):
    def __init__(self, requires_grad=True):
        super().__init__()
        self.requires_grad = requires_grad

    def forward(self, input1, input2):
        return DepthFlowProjectionLayer.apply(input1, input2, self.requires_grad)


class minDepthFlowProjectionLayer(Function):
    """MinDepthFlowProjection (minDepthFlowProjectionLayer.py:7-100), deterministic: the closest surface (largest
    input2) wins each cell; see include/vfidkr_b200.h.  Same signature and return values as the reference layer."""

    @staticmethod
    def forward(ctx, input1, input2, requires_grad):
        check_input(input1, "input1")
        check_input(input2, "input2")
        B, ch, H, W = input1.shape
        if ch != 2:
            raise _lib.VfidkrError("input1 must be a [B,2,H,W] flow")
        if input2.shape != (B, 1, H, W):
            raise _lib.VfidkrError(f"input2 must be [B,1,H,W] = {(B, 1, H, W)}, got {tuple(input2.shape)}")
        fillhole = 1 if requires_grad == False else 0   # noqa: E712  (minDepthFlowProjectionLayer.py:19)
        count = torch.empty((B, 1, H, W), dtype=input1.dtype, device=input1.device)
        output = torch.empty_like(input1)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_mindepthflowprojection_forward", ptr(input1), ptr(input2), ptr(count), ptr(output),
                      B, H, W, fillhole, stream_ptr(input1.device))
        ctx.save_for_backward(input1, input2, count)
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        input1, input2, count = ctx.saved_tensors
        gradoutput = gradoutput.contiguous()
        B, _, H, W = input1.shape
        gradinput1, gradinput2 = torch.empty_like(input1), torch.empty_like(input2)
        with torch.cuda.device(input1.device):
            _lib.call("vfidkr_mindepthflowprojection_backward", ptr(input1), ptr(input2), ptr(count), ptr(gradoutput),
                      ptr(gradinput1), ptr(gradinput2), B, H, W, stream_ptr(input1.device))
        return gradinput1, gradinput2, None   # minDepthFlowProjectionLayer.py:100


class minDepthFlowProjectionModule(Module):
    def __init__(self, requires_grad=True):
        super().__init__()
        self.requires_grad = requires_grad

    def forward(self, input1, input2):
        return minDepthFlowProjectionLayer.apply(input1, input2, self.requires_grad)


def flow_upsample4(flow_lowres, scale0=1.0, scale1=1.0):
    """Upsample(scale_factor=4, mode='bilinear')(scale0 * flow_lowres * scale1) (networks/DAIN.py:306-308) as one kernel."""
    check_input(flow_lowres, "flow_lowres")
    B, ch, h, w = flow_lowres.shape
    out = torch.empty((B, ch, 4 * h, 4 * w), dtype=flow_lowres.dtype, device=flow_lowres.device)
    if ch != 2:
        raise _lib.VfidkrError("flow_lowres must be [B,2,h,w]")
    with torch.cuda.device(flow_lowres.device):
        _lib.call("vfidkr_flow_upsample4", ptr(flow_lowres), float(scale0), float(scale1), ptr(out), B, h, w,
                  stream_ptr(flow_lowres.device))
    return out


class _FlowProjectLowresLayer(Function):
    """(Depth)FlowProjection fed by the quarter-resolution flow, differentiable: the forward never
    materialises the enlarged flow; the backward enlarges it once (flow_upsample4), runs the projection's own backward
    on it (flowprojection_cuda_kernel.cu:266-297 / depthflowprojection_cuda_kernel.cu:276-337 -- which, as in the
    reference, ignores hole filling) and folds the result back to the low resolution with the adjoint of the
    enlargement -- the chain autograd would build for networks/DAIN.py:306-308 + FlowProject."""

    @staticmethod
    def forward(ctx, flow_lowres, depth, scale0, scale1, fillhole):
        B, _, h, w = flow_lowres.shape
        if fillhole and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            # the backward needs the averaged output BEFORE hole filling (the reference never fills holes in training)
            raise _lib.VfidkrError("flow_project_lowres: hole filling is an inference feature (FlowProjectionLayer.py:23); "
                                   "pass fillhole=False when a gradient is required")
        count = torch.empty((B, 1, 4 * h, 4 * w), dtype=flow_lowres.dtype, device=flow_lowres.device)
        output = torch.empty((B, 2, 4 * h, 4 * w), dtype=flow_lowres.dtype, device=flow_lowres.device)
        with torch.cuda.device(flow_lowres.device):
            _lib.call("vfidkr_flowprojection_forward_lowres", ptr(flow_lowres), float(scale0), float(scale1),
                      ptr(depth) if depth is not None else None, ptr(count), ptr(output), B, h, w, 1 if fillhole else 0,
                      stream_ptr(flow_lowres.device))
        ctx.save_for_backward(flow_lowres, depth, count, output)
        ctx.scales = (float(scale0), float(scale1))
        return output

    @staticmethod
    def backward(ctx, gradoutput):
        flow_lowres, depth, count, output = ctx.saved_tensors
        s0, s1 = ctx.scales
        B, _, h, w = flow_lowres.shape
        H, W = 4 * h, 4 * w
        gradoutput = gradoutput.contiguous()
        full = flow_upsample4(flow_lowres, s0, s1)
        g_full = torch.empty_like(full)
        g_depth = torch.empty_like(depth) if depth is not None else None
        g_low = torch.empty_like(flow_lowres)
        sp = stream_ptr(flow_lowres.device)
        with torch.cuda.device(flow_lowres.device):
            if depth is None:
                _lib.call("vfidkr_flowprojection_backward", ptr(full), ptr(count), ptr(gradoutput), ptr(g_full), B, H, W, sp)
            else:
                _lib.call("vfidkr_depthflowprojection_backward", ptr(full), ptr(depth), ptr(count), ptr(output), ptr(gradoutput),
                          ptr(g_full), ptr(g_depth), B, H, W, sp)
            _lib.call("vfidkr_flow_upsample4_backward", ptr(g_full), s0, s1, ptr(g_low), B, h, w, sp)
        return g_low, g_depth, None, None, None


def flow_project_lowres(flow_lowres, scale0=1.0, scale1=1.0, depth=None, fillhole=True):
    """(Depth)FlowProjection of the x4-enlarged, scaled flow without materialising it.
    Equals FlowProjectionModule(not fillhole)(flow_upsample4(flow_lowres, scale0, scale1)) -- DepthFlowProjection with
    `depth` -- including its gradients with respect to flow_lowres and depth (fillhole=False, as in training)."""
    check_input(flow_lowres, "flow_lowres")
    B, ch, h, w = flow_lowres.shape
    if ch != 2:
        raise _lib.VfidkrError("flow_lowres must be [B,2,h,w]")
    if depth is not None:
        check_input(depth, "depth")
        if depth.shape != (B, 1, 4 * h, 4 * w):
            raise _lib.VfidkrError(f"depth must be [B,1,4h,4w] = {(B, 1, 4 * h, 4 * w)}")
    return _FlowProjectLowresLayer.apply(flow_lowres, depth, scale0, scale1, fillhole)
