"""Distortion masks for the segment-wise beamformer (tssep/train/enhancer_distortion_mask.py:9-55).

``masks`` (1, speakers, ...) -> (2, speakers, ...): the second plane is the weight of everything that is NOT the
speaker.  numpy in -> numpy out (host arithmetic on a mask-sized array, as the reference); CUDA tensor in -> CUDA tensor.
"""
from __future__ import annotations

import numpy as np
import torch


class OneMinus:
    """noise mask = max(1 - mask, 0) (enhancer_distortion_mask.py:9-21)."""

    def __call__(self, masks):
        assert masks.shape[0] == 1, masks.shape
        if isinstance(masks, torch.Tensor):
            return torch.cat([masks, torch.clamp(1 - masks, min=0)], dim=0)
        return np.concatenate([masks, np.maximum(1 - masks, 0)], axis=0)


class SumCrossTalker:
    """noise mask of a speaker = sum of the masks of all other speakers, at least ``eps``
    (enhancer_distortion_mask.py:24-55)."""

    def __init__(self, eps=0.0001):
        self.eps = eps

    def __call__(self, masks):
        assert masks.shape[0] == 1, masks.shape
        if isinstance(masks, torch.Tensor):
            total = masks.sum(dim=1, keepdim=True)
            return torch.cat([masks, torch.clamp(total - masks, min=self.eps)], dim=0)
        speakers = masks.shape[1]
        # summed speaker by speaker like the reference (total - own would round differently)
        noise = np.stack([np.sum(np.delete(masks, spk, axis=1), axis=1) for spk in range(speakers)], axis=1)
        return np.concatenate([masks, np.maximum(noise, self.eps)], axis=0)
