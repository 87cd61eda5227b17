"""Small helpers shared by the operator front-ends (device pointers, stream, argument checks)."""
from __future__ import annotations

from ctypes import c_void_p

import torch


def check_input(t: torch.Tensor, name: str) -> None:
    """Mirror of the reference's pre-conditions: CUDA, float32, contiguous
    (`assert(input1.is_contiguous())`, FilterInterpolationLayer.py:16-18).  No CPU path exists."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: vfidkr_b200 has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    if not t.is_contiguous():
        # the reference asserts; an assert disappears under `python -O`, and the C ABI takes dense NCHW pointers
        raise ValueError(f"{name} must be contiguous (dense NCHW): call .contiguous() first")


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
