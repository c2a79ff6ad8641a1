"""Scalar heads needed to instantiate the reference configs (tssep/train/loss.py).

Only the pieces the inference path touches are provided: ``LogMAE``
(loss.py:219-247) for the end-to-end golden / training-forward config and a
minimal ``VADSigmoidBCE`` (loss.py:272-345) so ``init_cfg_tsvad.yaml`` loads.
They are a few torch ops on already-computed outputs, not part of the hot path.
"""
from __future__ import annotations

import torch

from .configurable import Configurable


class LogMAE(Configurable):
    name = "LogMAE"

    def __init__(self, target: str = "speaker_reverberation_early_ch0", pit: bool = False):
        if pit:
            raise NotImplementedError("pit=True")
        self.target, self.pit = target, pit

    def targets(self, lower=False):
        return [self.target.lower() if lower else self.target]

    def __call__(self, estimate, target):
        return torch.log10((estimate - target).abs().mean(dim=-1).sum(dim=-1))

    def from_ex_out(self, ex, out, model=None, summary=None):
        return self(out.time_estimate, ex[self.target])


class VADSigmoidBCE(Configurable):
    name = "VADSigmoidBCE"

    def __init__(self, target: str = "Vad", pit: bool = False, magnitude_threshold: float = 0.05):
        if pit:
            raise NotImplementedError("pit=True")
        self.target, self.pit, self.magnitude_threshold = target, pit, magnitude_threshold

    def targets(self, lower=False):
        return [self.target.lower() if lower else self.target]

    def __call__(self, logit, target):
        """logit (..., K, 1, T, F) -> frequency mean (loss.py:343) -> BCE with logits vs (…, K, T)."""
        vad_logit = logit.mean(dim=-1).squeeze(-2)
        return torch.nn.functional.binary_cross_entropy_with_logits(vad_logit, target.to(vad_logit.dtype))
