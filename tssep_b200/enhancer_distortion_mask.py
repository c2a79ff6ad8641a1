"""Distortion masks for the segment-wise beamformer (tssep/train/enhancer_distortion_mask.py:9-55).

``masks`` (1, speakers, ...) -> (2, speakers, ...): the second plane is the weight of everything that is NOT the
speaker.  CUDA tensor in -> CUDA tensor out; numpy in -> moved to the current CUDA device, numpy out (no CPU path).
"""
from __future__ import annotations

import numpy as np
import torch


def _to_cuda(masks):
    if isinstance(masks, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("tssep_b200 needs a CUDA device (no CPU fallback)")
        return torch.as_tensor(masks).cuda(), True
    if not masks.is_cuda:
        raise RuntimeError(f"tssep_b200 operators run on CUDA tensors only (no CPU fallback); got a tensor on {masks.device}")
    return masks, False


class OneMinus:
    """noise mask = max(1 - mask, 0) (enhancer_distortion_mask.py:9-21)."""

    def __call__(self, masks):
        assert masks.shape[0] == 1, masks.shape
        m, was_np = _to_cuda(masks)
        out = torch.cat([m, torch.clamp(1 - m, min=0)], dim=0)
        return out.cpu().numpy() if was_np else out


class SumCrossTalker:
    """noise mask of a speaker = sum of the masks of all other speakers, at least ``eps``
    (enhancer_distortion_mask.py:24-55)."""

    def __init__(self, eps=0.0001):
        self.eps = eps

    def __call__(self, masks):
        assert masks.shape[0] == 1, masks.shape
        m, was_np = _to_cuda(masks)
        speakers = m.shape[1]
        # summed speaker by speaker in the reference's order (total - own would round differently)
        noise = torch.stack([m[:, [j for j in range(speakers) if j != spk]].sum(dim=1) for spk in range(speakers)], dim=1)
        out = torch.cat([m, torch.clamp(noise, min=self.eps)], dim=0)
        return out.cpu().numpy() if was_np else out
