"""Diarization post-processing on the device: frame activity (frequency mean of the
mask), running-median smoothing, thresholding and run-length segment extraction.

The reference has no implementation of this stage (it lives in fgnt/tssep_data;
anchors: tssep/util/utils.py:11-129, tssep/train/loss.py:343); the behaviour is
specified by ``oracle/tssep_oracle.py::diarize_reference`` (parity unpinned).
"""
from __future__ import annotations

import dataclasses

import torch

from . import _lib, torch_ops


@dataclasses.dataclass
class Diarization:
    activity: torch.Tensor   # (..., K, T) float32  mean_f mask
    smooth: torch.Tensor     # (..., K, T) float32  running median
    active: torch.Tensor     # (..., K, T) uint8
    segments: torch.Tensor   # (..., K, max_segments, 2) int32 sample intervals [start, end)
    counts: torch.Tensor     # (..., K) int32 number of runs (may exceed max_segments)

    def to_lists(self):
        """Host copy: list (per leading index) of lists of (start, end) tuples."""
        seg = self.segments.cpu().numpy().reshape(-1, *self.segments.shape[-2:])
        cnt = self.counts.cpu().numpy().reshape(-1)
        return [[(int(a), int(b)) for a, b in s[: min(c, s.shape[0])]] for s, c in zip(seg, cnt)]


def diarize(mask: torch.Tensor, fe, *, num_samples=None, threshold=0.5, median_width=1, max_segments=256,
            activity: torch.Tensor = None) -> Diarization:
    """mask (..., K, 1, T, F) float32 on the device.  ``activity`` (..., K, T): the frame activity when the enhancement
    kernel already produced it (``Masking.apply(activity_out=...)``); the mask is then not read again."""
    _lib.require_cuda(mask, activity)
    if mask.shape[-3] != 1:
        raise ValueError(f"expected nmask == 1, got {tuple(mask.shape)}")
    lead = mask.shape[:-3]
    T, F = mask.shape[-2:]
    n = mask.numel() // (T * F)
    dev = mask.device
    if activity is None:
        from .ops import row_pitch_view

        m, mask_pitch = row_pitch_view(mask)
        act = torch.empty((*lead, T), dtype=torch.float32, device=dev)
        torch_ops.op.activity(m, n, T, F, mask_pitch, act)
    else:
        if tuple(activity.shape) != (*lead, T) or activity.dtype != torch.float32 or not activity.is_contiguous():
            raise ValueError(f"activity must be a contiguous float32 tensor of shape {(*lead, T)}")
        act = activity
    smooth = torch.empty_like(act)
    active = torch.empty((*lead, T), dtype=torch.uint8, device=dev)
    seg = torch.zeros((*lead, max_segments, 2), dtype=torch.int32, device=dev)
    cnt = torch.empty(lead, dtype=torch.int32, device=dev)
    torch_ops.op.median_threshold(act, n, T, int(median_width), float(threshold), smooth, active)
    torch_ops.op.segments(active, n, T, fe.window_length, fe.shift, int(bool(fe.fading)),
                          -1 if num_samples is None else int(num_samples), seg, cnt, max_segments, 0)
    return Diarization(act, smooth, active, seg, cnt)
