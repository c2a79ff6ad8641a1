"""Drop-in ``Model`` glue for the inference path (tssep/train/model.py:70-164, :454-536, :661-664).

``Model.forward(ex)`` keeps the reference's call contract (same ``ex`` keys in,
``ForwardOutput`` out, ``ex['Observation']`` / ``ex['Input']`` / ``ex['AuxInput']``
filled in).  ``Model.separate`` is the throughput entry point: a batch of
independent meetings in, every output of the path (mask, logit, stft_estimate,
time_estimate and optionally diarization segments) out, all on the device.
Dataset plumbing, review summaries and the trainer are out of scope.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from . import _lib
from .configurable import Configurable, import_class
from .enhancer import Masking
from .feature_extractor import _compute_features


@dataclasses.dataclass
class ForwardOutput:
    mask: torch.Tensor
    logit: torch.Tensor
    embedding: torch.Tensor = None
    stft_estimate: torch.Tensor = None
    time_estimate: torch.Tensor = None

    vad_mask: torch.Tensor = None
    vad_logit: torch.Tensor = None


class Model(Configurable, torch.nn.Module):
    ForwardOutput = ForwardOutput

    @classmethod
    def finalize_dogmatic_config(cls, config):
        # same defaults as tssep/train/model.py:119-149
        config["fe"] = dict(factory="tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT", size=1024, shift=256,
                            window="hann")
        config["reader"] = dict(factory="tssep_b200.data.DummyReader")
        config["enhancer"] = dict(factory="tssep_b200.enhancer.Masking")
        fe = Configurable.from_config(config["fe"])
        is_masking = issubclass(import_class(config["enhancer"]["factory"]), Masking)
        config["mask_estimator"] = dict(factory="tssep_b200.net.MaskEstimator_v2", idim=fe.output_size,
                                        odim=fe.frequencies, nmask=1 if is_masking else 2)
        config["loss"] = dict(factory="tssep_b200.loss.LogMAE")

    def __init__(self, fe=None, reader=None, mask_estimator=None, enhancer=None, loss=None):
        super().__init__()
        self.fe = fe
        self.reader = reader
        self.mask_estimator = mask_estimator
        self.enhancer = enhancer
        self.loss = loss

    # -- helpers -----------------------------------------------------------------
    def example_to_device(self, ex, device):
        """numpy / CPU example -> tensors on ``device`` (tssep/train/model.py:166-180)."""
        out = dict(ex)
        if "audio_data" in ex:
            out["observation"] = ex["audio_data"]["observation"]
            for k, v in ex["audio_data"].items():
                if k != "observation":
                    out.setdefault(k, v)
        out.setdefault("reference_channel", 0)
        targets = self.loss.targets() + self.loss.targets(lower=True) if self.loss is not None else []
        for k in ["observation", "auxInput", "Input", *targets]:
            if k in out and not isinstance(out[k], (str, int)):
                out[k] = torch.as_tensor(np.asarray(out[k]) if not torch.is_tensor(out[k]) else out[k]).to(device)
        return out

    def _features(self, ex, couple=None):
        ref = ex["reference_channel"]
        if not isinstance(ref, int):
            raise NotImplementedError(type(ref), ref)
        if "Observation" not in ex:
            ex["Observation"] = self.fe.stft(ex["observation"])
        Xr = ex["Observation"][..., ref, :, :]
        feats = _compute_features(self.fe, Xr, want_f32=True, want_bf16=True, couple=couple)
        return feats

    # -- reference contract --------------------------------------------------------
    def forward(self, ex, feature_transform=None, with_time_estimate=False) -> ForwardOutput:
        """tssep/train/model.py:465-536.  ``with_time_estimate=True`` additionally runs the iSTFT of
        ``Model.review`` (model.py:661-664) in the same enhancement kernel."""
        ex["AuxInput"] = [a for a in ex["auxInput"]]
        if self.training and torch.is_grad_enabled():
            return self._forward_train(ex, feature_transform)
        bf16 = None
        if "Input" not in ex:
            feats = self._features(ex)
            ex["Input"] = feats["f32"]
            bf16 = (feats["bf16"], feats["ld"])
        if feature_transform is not None:
            ex["Input"] = feature_transform(ex["Input"])
            bf16 = None
        ex = self.reader.data_hooks.pre_net(ex) if self.reader is not None else ex
        me_out = self.mask_estimator(ex["Input"], ex["AuxInput"], _features_bf16=bf16)
        stft_estimate = time_estimate = None
        if "Observation" in ex:
            if isinstance(self.enhancer, Masking):
                n = ex["observation"].shape[-1] if "observation" in ex else None
                stft_estimate, time_estimate = Masking.apply(
                    me_out.mask, ex["Observation"], ex["reference_channel"], self.fe, want_estimate=True,
                    want_time=with_time_estimate and n is not None, num_samples=n)
            else:
                stft_estimate = self.enhancer(me_out.mask, ex, self)
        return ForwardOutput(mask=me_out.mask, logit=me_out.logit, vad_mask=me_out.vad_mask,
                             vad_logit=me_out.vad_logit, embedding=me_out.embedding, stft_estimate=stft_estimate,
                             time_estimate=time_estimate)

    def _forward_train(self, ex, feature_transform=None) -> ForwardOutput:
        """Training step forward (BASELINE config 5): the same path with autograd through it -- features (inputs, no
        gradient), mask estimator (``MaskEstimator_v2.forward_train``), Masking as a torch product, and the iSTFT of
        ``Model.review`` (model.py:661-664) through ``tssep_b200.autograd.ISTFTFn`` so that ``LogMAE`` on
        ``time_estimate`` back-propagates."""
        from .autograd import ISTFTFn

        if "Input" not in ex:
            with torch.no_grad():
                ex["Input"] = self._features(ex)["f32"]
        if feature_transform is not None:
            ex["Input"] = feature_transform(ex["Input"])
        ex = self.reader.data_hooks.pre_net(ex) if self.reader is not None else ex
        me_out = self.mask_estimator(ex["Input"], ex["AuxInput"])
        stft_estimate = time_estimate = None
        if "Observation" in ex:
            if not isinstance(self.enhancer, Masking):
                raise NotImplementedError("training through enhancers other than Masking")
            obs = ex["Observation"][..., ex["reference_channel"], :, :]
            stft_estimate = obs[..., None, :, :] * torch.squeeze(me_out.mask, dim=-3)   # enhancer.py:73-100
            if "observation" in ex:
                time_estimate = ISTFTFn.apply(stft_estimate, self.fe, ex["observation"].shape[-1])
        return ForwardOutput(mask=me_out.mask, logit=me_out.logit, vad_mask=me_out.vad_mask, vad_logit=me_out.vad_logit,
                             embedding=me_out.embedding, stft_estimate=stft_estimate, time_estimate=time_estimate)

    def istft(self, out: ForwardOutput, num_samples: Optional[int]):
        """The time-domain step of ``Model.review`` (tssep/train/model.py:661-664)."""
        out.time_estimate = self.fe.istft(out.stft_estimate, num_samples=num_samples)
        return out

    # -- throughput entry point ------------------------------------------------------
    @torch.no_grad()
    def separate(self, observation: torch.Tensor, aux: torch.Tensor, *, want_estimate=True, want_time=True,
                 diarize: Optional[dict] = None) -> ForwardOutput:
        """A batch of independent single-channel meetings.

        observation (M, N) or (M, 1, N) float32, aux (M, K, A) float32, both on the device.  Each
        meeting is processed exactly as a separate unbatched ``forward`` call would (own feature
        statistics, one ``np.random.permutation`` draw per meeting, in order).
        ``diarize`` = dict(threshold=, median_width=, max_segments=) adds ``.segments``.
        """
        (_, _, out), = self.separate_waves(observation, aux, wave=None, want_estimate=want_estimate, want_time=want_time,
                                           diarize=diarize)
        return out

    @torch.no_grad()
    def separate_waves(self, observation: torch.Tensor, aux: torch.Tensor, *, wave: Optional[int] = None,
                       out_wave: Optional[int] = None, want_estimate=True, want_time=True,
                       diarize: Optional[dict] = None, time_out: Optional[torch.Tensor] = None):
        """``separate`` for more meetings than one wave of the recurrence kernels holds: a generator of
        ``(lo, hi, ForwardOutput)`` for meetings [lo, hi), ``wave`` meetings at a time.

        Front end, pre_net and the speaker-concat layer run once for all M meetings (their recurrences
        cost the same T dependent steps for 1 or 100 rows), the K-rows-per-meeting layers and every
        output stage run per wave (``MaskEstimator_v2.forward_waves``).  A consumer that drops a wave's
        big outputs (mask, logit, stft_estimate: 2.5 GB per 10-min meeting) before asking for the next
        one lets them share memory; ``out_wave`` (default ``wave``) is the number of meetings per yielded
        output.  ``time_out`` (M, K, N) float32: optional caller-owned buffer the separated signals are written
        to (each yielded ``time_estimate`` is then a view of it); a callable is invoked right before the first
        output wave is written and must return the buffer (a serving loop waits there for the previous batch's
        device-to-host copy).  Results do not depend on ``wave`` / ``out_wave``.
        """
        _lib.require_cuda(observation, aux)
        if observation.dim() == 2:
            observation = observation[:, None, :]
        n = observation.shape[-1]
        ex = {"observation": observation, "reference_channel": 0}
        feats = self._features(ex, couple=False)
        Obs = ex["Observation"]
        waves = self.mask_estimator.forward_waves(feats["f32"], [a for a in aux], wave=wave, out_wave=out_wave,
                                                  _features_bf16=(feats["bf16"], feats["ld"]))
        del feats
        for lo, hi, me_out in waves:
            if callable(time_out):  # resolved when the first output wave is about to be written
                time_out = time_out()
            # (Masking.apply can also reduce the mask rows it reads to the frame activity of the diarization stage; measured
            # on B200 the stand-alone reduction at HBM speed is cheaper: 2.73 + 0.39 ms vs 3.28 ms per 4 meetings, because
            # the transform kernel has no registers to spare -- profiles/r2_istft_activity_ab.txt)
            est, time = Masking.apply(me_out.mask, Obs[lo:hi], 0, self.fe, want_estimate=want_estimate,
                                      want_time=want_time, num_samples=n,
                                      time_out=None if time_out is None else time_out[lo:hi])
            out = ForwardOutput(mask=me_out.mask, logit=me_out.logit, embedding=me_out.embedding, stft_estimate=est,
                                time_estimate=time, vad_mask=me_out.vad_mask, vad_logit=me_out.vad_logit)
            if diarize is not None:
                from .postprocess import diarize as run_diarize

                out.segments = run_diarize(me_out.mask, self.fe, num_samples=n, **diarize)
            del me_out, est, time
            yield lo, hi, out
            del out
