"""Losses of the reference (tssep/train/loss.py): ``LogMAE`` on the separated signals (TS-SEP) and ``VADSigmoidBCE`` on
the frequency-averaged logits (TS-VAD).  A few torch ops on already-computed outputs -- not part of the hot path; in
the training step (BASELINE config 5) autograd starts here and runs back through ``tssep_b200.autograd``.
"""
from __future__ import annotations

import numpy as np
import torch

from .configurable import Configurable


class _Loss(Configurable):
    def __init__(self, target: str, pit: bool = False):
        if pit:
            raise NotImplementedError("pit=True (padertorch pit_loss) is not implemented")
        self.target, self.pit = target, pit

    @property
    def name(self):
        return type(self).__name__

    def targets(self, lower=False):
        return [self.target.lower() if lower else self.target]

    def __call__(self, estimate, target):
        assert estimate.shape == target.shape, (estimate.shape, target.shape)
        return self.loss_fn(estimate, target)


class MAE(_Loss):
    """tssep/train/loss.py:192-216."""

    def __init__(self, target: str = "speaker_reverberation_early_ch0", pit: bool = False):
        super().__init__(target, pit)

    def loss_fn(self, estimate, target):
        return (estimate - target).abs().mean(dim=-1).sum(dim=-1)

    def from_ex_out(self, ex, out, model=None, summary=None):
        return self(out.time_estimate, ex[self.target])


class LogMAE(MAE):
    """``log10(sum_k mean_n |est - tgt|)`` per item (tssep/train/loss.py:219-247).

    >>> _ = torch.manual_seed(0)
    >>> target = torch.rand((2, 10000))
    >>> estimate = target + 0.5 * torch.rand((2, 10000))
    >>> LogMAE(pit=False)(estimate, target)
    tensor(-0.2995)
    """

    def loss_fn(self, estimate, target):
        return torch.log10(super().loss_fn(estimate, target))


class VADSigmoidBCE(_Loss):
    """Binary cross entropy between the frequency mean of the logits and the frame activity
    (tssep/train/loss.py:272-345).

    >>> _ = torch.manual_seed(0)
    >>> target = torch.rand((2, 100, 257))
    >>> estimate = target + 0.5 * torch.rand((2, 100, 257))
    >>> VADSigmoidBCE(pit=False, target='Speaker_reverberation_early')(estimate, target)
    tensor(0.3867)
    """

    def __init__(self, target: str = "Vad", pit: bool = False, magnitude_threshold: float = 0.05):
        super().__init__(target, pit)
        assert 0 < magnitude_threshold < 1, magnitude_threshold
        self.magnitude_threshold = magnitude_threshold

    def loss_fn(self, estimate, target):
        return torch.nn.functional.binary_cross_entropy_with_logits(estimate, target, reduction="none").mean(dim=(-1, -2))

    def prepare_target(self, target, dtype=None):
        if self.target in ["vad", "Vad"]:
            return target
        if isinstance(target, torch.Tensor):
            dtype = target.real.dtype if dtype is None else dtype
            t = target.abs().sum(dim=-1)
            t = t / torch.amax(t, dim=-1, keepdim=True)
            return (t > self.magnitude_threshold).type(dtype)
        dtype = target.real.dtype if dtype is None else dtype
        t = np.abs(target).sum(axis=-1)
        t = t / np.amax(t, axis=-1, keepdims=True)
        return (t > self.magnitude_threshold).astype(dtype)

    def __call__(self, estimate: torch.Tensor, target):
        if not isinstance(target, torch.Tensor):
            target = torch.stack(list(target))
        if self.target not in ["vad", "Vad"]:
            assert estimate.shape == target.shape, (estimate.shape, target.shape)
            assert estimate.ndim > 2, estimate.shape
            target = self.prepare_target(target)
        estimate = torch.mean(estimate, dim=-1)  # loss.py:343
        assert estimate.shape == target.shape, (estimate.shape, target.shape)
        return self.loss_fn(estimate, target.to(estimate.dtype))

    def from_ex_out(self, ex, out, model, summary=None):
        """tssep/train/loss.py:117-146: logits without the mask axis vs ``ex['Vad']`` (made from the sample activity with
        ``stft_vad`` when absent)."""
        from .util.utils import stft_vad

        estimate = torch.squeeze(out.logit, dim=-3)
        if self.target not in ex:
            if self.target == "Vad":
                ex[self.target] = stft_vad(ex[self.target.lower()], model.fe.window_length, model.fe.shift, model.fe.fading)
            else:
                ex[self.target] = model.fe.stft(ex[self.target.lower()])
        return self(estimate, ex[self.target])
