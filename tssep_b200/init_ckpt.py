"""Checkpoint initialisers of the reference's two-stage recipe (tssep/train/init_ckpt.py:15-88): a TS-SEP model starts
from a trained TS-VAD model whose head ``linear2`` has one output per (speaker, mask) and is broadcast along frequency.
They take the model (the reference passes its ``Experiment``; only ``eg.trainer.model`` is used there)."""
from __future__ import annotations

import dataclasses
from pathlib import Path

import torch

from .configurable import Configurable


@dataclasses.dataclass
class InitCheckPoint(Configurable):
    init_ckpt: "str | Path" = None
    strict: bool = True

    def _state(self, ckpt):
        if isinstance(ckpt, dict):
            return ckpt
        ckpt = Path(ckpt)
        assert ckpt.exists(), ckpt
        return torch.load(str(ckpt), map_location="cpu")

    def load_model_state_dict(self, model: torch.nn.Module, ckpt):
        return model.load_state_dict(self._state(ckpt)["model"], strict=self.strict)

    def __call__(self, model: torch.nn.Module):
        if self.init_ckpt is not None:
            return self.load_model_state_dict(model, self.init_ckpt)


@dataclasses.dataclass
class InitCheckPointVAD2Sep(InitCheckPoint):
    """Broadcast of the last layer (init_ckpt.py:39-88): ``repeat`` = ``torch.repeat_interleave`` along every axis that
    is smaller in the checkpoint, which relies on the ``(spk mask freq)`` speaker-major row order of ``linear2``."""

    bcast: tuple = ("mask_estimator.post_net.linear2.weight", "mask_estimator.post_net.linear2.bias")
    mode: str = "repeat"

    def load_model_state_dict(self, model: torch.nn.Module, ckpt):
        state = dict(self._state(ckpt)["model"])
        for k in self.bcast:
            shape = model.get_parameter(k).shape
            p = state[k]
            assert len(p.shape) == len(shape), (p.shape, shape)
            assert self.mode == "repeat", f"mode {self.mode!r} is not implemented (neither in the reference)"
            for i, (actual, desired) in enumerate(zip(p.shape, shape)):
                if actual == desired:
                    continue
                if actual > desired or desired % actual != 0:
                    raise ValueError((p.shape, shape, actual, desired))
                p = torch.repeat_interleave(p, desired // actual, dim=i)
            state[k] = p
        return model.load_state_dict(state, strict=self.strict)
