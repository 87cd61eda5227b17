"""Frame I/O boundary of the demo drivers as device kernels.

demo_MiddleBury.py:276-364 (and colab_interpolate.py:85-148) turn uint8 HWC frames into float CHW / 255, pad them by
replication to the next multiple of 128 (or by 32 + 32 when a dimension already is one), run the network, crop, scale
by 255, clip, round and convert back to uint8 -- in numpy and torch on the host.  Here both directions are one kernel
each on the device, so a streaming caller moves uint8 frames over PCIe (a quarter of the float bytes).
Bit-exact against the numpy formulation.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._common import stream_ptr


def frame_padding(size: int) -> tuple[int, int]:
    """(leading pad, padded size) of one dimension, demo_MiddleBury.py:286-301."""
    padded = ctypes.c_int(0)
    lead = _lib.load().vfidkr_frame_padding(int(size), ctypes.byref(padded))
    return int(lead), int(padded.value)


def frames_to_padded(frames: torch.Tensor) -> torch.Tensor:
    """uint8 [B,H,W,3] on a CUDA device -> float32 [B,3,Hp,Wp] in [0,1], replication-padded."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_cuda or not frames.is_contiguous():
        raise _lib.VfidkrError("frames must be a contiguous uint8 [B,H,W,3] CUDA tensor (there is no CPU path)")
    B, H, W, _ = frames.shape
    out = torch.empty((B, 3, frame_padding(H)[1], frame_padding(W)[1]), dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        _lib.call("vfidkr_frames_u8_to_padded_f32", frames.data_ptr(), out.data_ptr(), B, H, W, stream_ptr(frames.device))
    return out


def padded_to_frames(padded: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """float32 [B,3,Hp,Wp] -> uint8 [B,height,width,3]: crop, 255 * clip(v, 0, 1), round half to even."""
    B = padded.shape[0]
    if padded.dtype != torch.float32 or not padded.is_cuda or not padded.is_contiguous() or \
            tuple(padded.shape) != (B, 3, frame_padding(height)[1], frame_padding(width)[1]):
        raise _lib.VfidkrError("padded must be a contiguous float32 [B,3,Hp,Wp] CUDA tensor with the padding of (height, width)")
    out = torch.empty((B, height, width, 3), dtype=torch.uint8, device=padded.device)
    with torch.cuda.device(padded.device):
        _lib.call("vfidkr_padded_f32_to_frames_u8", padded.data_ptr(), out.data_ptr(), B, height, width, stream_ptr(padded.device))
    return out
