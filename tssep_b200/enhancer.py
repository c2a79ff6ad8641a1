"""Drop-in ``Masking`` enhancer (tssep/train/enhancer.py:21-32, :73-100).

``stft_estimate = Observation[..., ref, None, :, :] * squeeze(mask, -3)`` computed by
``tssep_mask_istft``; ``apply`` can also return the iSTFT of the product from the
same kernel launch (mask x X -> irfft -> synthesis window -> overlap-add).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, torch_ops
from .configurable import Configurable


class ABC(Configurable, torch.nn.Module):
    @property
    def name(self):
        return self.__class__.__name__

    def __call__(self, masks, ex, model):
        raise NotImplementedError()


class Masking(ABC):
    def __call__(self, masks: torch.Tensor, ex, model):
        return self.apply(masks, ex["Observation"], ex["reference_channel"], fe=model.fe)[0]

    @staticmethod
    def apply(masks, Observation, reference_channel, fe, want_estimate=True, want_time=False, num_samples=None,
              time_out=None, activity_out=None):
        """masks (K,1,T,F) or (B,K,1,T,F) f32; Observation ([B,] C, T, F) complex64 (or without the channel
        axis when ``reference_channel`` is None).  Returns (stft_estimate | None, time_estimate | None).
        ``time_out``: optional preallocated contiguous float32 device tensor ([B,] K, n) for time_estimate (a
        serving loop that ships the audio to the host keeps its own buffers instead of churning the allocator).
        ``activity_out``: optional float32 device tensor ([B,] K, T) that receives the frame activity
        ``mean_f mask`` (the first stage of ``postprocess.diarize``), reduced from the mask rows the kernel reads anyway."""
        batched = {4: False, 5: True}[masks.dim()]
        if isinstance(Observation, np.ndarray):
            Observation = torch.as_tensor(Observation, device=masks.device)
        _lib.require_cuda(masks, Observation)
        if reference_channel is None:
            assert Observation.dim() == (3 if batched else 2), Observation.shape
            obs = Observation
        else:
            assert Observation.dim() == (4 if batched else 3), Observation.shape
            obs = Observation[..., reference_channel, :, :]
        if masks.shape[-3] != 1:
            raise ValueError(f"Masking needs nmask == 1, got mask shape {tuple(masks.shape)}")
        obs = obs.to(torch.complex64).contiguous()
        from .ops import row_pitch_view

        m, mask_pitch = row_pitch_view(masks)   # the head GEMM's padded rows are read in place
        T, F = m.shape[-2:]
        K = m.shape[-4]
        Z = m.shape[0] if batched else 1
        assert obs.shape[-2:] == (T, F), (obs.shape, m.shape)
        lead = (Z, K) if batched else (K,)
        est = torch.empty((*lead, T, F), dtype=torch.complex64, device=m.device) if want_estimate else None
        time = None
        if want_time:
            total = (T - 1) * fe.shift + fe.window_length - (2 * (fe.window_length - fe.shift) if fe.fading else 0)
            n = total if num_samples is None else min(int(num_samples), total)
            if time_out is not None:
                if (tuple(time_out.shape) != (*lead, n) or time_out.dtype != torch.float32 or not time_out.is_contiguous()
                        or time_out.device != m.device):
                    raise ValueError(f"time_out must be a contiguous float32 tensor of shape {(*lead, n)} on {m.device}")
                time = time_out
            else:
                time = torch.empty((*lead, n), dtype=torch.float32, device=m.device)
        if activity_out is not None:
            _lib.require_cuda(activity_out)
            if (tuple(activity_out.shape) != (*lead, T) or activity_out.dtype != torch.float32
                    or not activity_out.is_contiguous()):
                raise ValueError(f"activity_out must be a contiguous float32 tensor of shape {(*lead, T)}")
        tab = fe._device_tables(m.device)
        torch_ops.op.mask_istft(obs, T * F, m, mask_pitch, Z, K, T, fe.size, fe.shift, fe.window_length, int(bool(fe.fading)),
                                tab["synwin"], tab["twiddle"], est, time, time.shape[-1] if time is not None else 0,
                                activity_out)
        return est, time


class TorchBF(ABC):
    """Mask-based MVDR beamformer, Souden formulation (tssep/train/enhancer.py:140-283): spatial covariance matrices of
    target and interference from the masks, ``w = (Phi_i^-1 Phi_t)[:, ref] / trace(Phi_i^-1 Phi_t)``, ``enh = w^H Y``.

    Same constructor and call contract as the reference: ``masks`` (K, nmask, T, F) or (B, K, nmask, T, F) with
    ``nmask`` 1 (interference weight = 1 - mask) or 2 (target, interference); ``ex['Observation']`` ([B,] D, T, F)
    complex.  The covariance sums and the D x D solves run in float64 (the reference works in complex128 throughout),
    the beamformer is applied in float32; the result has the dtype of ``Observation``."""

    def __init__(self, bf="mvdr_souden", masking=False, masking_eps=0.0, eps=None):
        super().__init__()
        assert bf == "mvdr_souden", (bf, "Only mvdr_souden is implemented")
        self.bf, self.eps, self.masking, self.masking_eps = bf, eps, masking, masking_eps

    def __call__(self, masks, ex, model=None):
        batched = {4: False, 5: True}[masks.dim()]
        ref = ex["reference_channel"]
        Obs = ex["Observation"]
        if isinstance(Obs, np.ndarray):
            Obs = torch.as_tensor(Obs, device=masks.device)
        assert Obs.dim() == (4 if batched else 3), Obs.shape
        _lib.require_cuda(masks, Obs)
        out_dtype = Obs.dtype
        Y = Obs.to(torch.complex64).contiguous()
        m = masks.float().contiguous()
        if not batched:
            Y, m = Y[None], m[None]
        Z, D, T, F = Y.shape
        K, nmask = m.shape[1], m.shape[2]
        assert m.shape == (Z, K, nmask, T, F), (m.shape, Y.shape)
        planes = K * nmask + (1 if nmask == 1 else 0)
        psd = torch.empty((Z, planes, F, D * (D + 1) // 2, 2), dtype=torch.float64, device=Y.device)
        torch_ops.op.bf_psd(Y, m, Z, K, nmask, D, T, F, psd)
        w = torch.empty((Z, K, F, D), dtype=torch.complex64, device=Y.device)
        eps = float(torch.finfo(torch.float64).tiny) if self.eps is None else float(self.eps)
        torch_ops.op.bf_mvdr_souden(psd, Z, K, nmask, D, F, int(ref), eps, w)
        enh = torch.empty((Z, K, T, F), dtype=torch.complex64, device=Y.device)
        torch_ops.op.bf_apply(Y, w, m if self.masking else None, Z, K, nmask, D, T, F, float(self.masking_eps), enh)
        self.last_beamformer = w if batched else w[0]
        enh = enh.to(out_dtype) if out_dtype.is_complex else enh
        return enh if batched else enh[0]


class WPE:
    """Weighted-prediction-error dereverberation in front of the feature extraction
    (tssep/train/enhancer.py:292-345; same constructor and call signature).

    ``Observation`` (D, T, F) complex, numpy (moved to the current CUDA device, result returned as numpy of the input's
    dtype) or a CUDA tensor.  One ``tssep_wpe`` call: statistics / solve / filter kernels of csrc/wpe.cu -- complex64
    signals, f64 accumulation of the correlation matrices and an f64 solve.  The arithmetic the reference delegates to
    ``nara_wpe`` (absent package) is restated from its publication: parity is pinned to the repo's own oracle only
    (oracle/tssep_oracle.py::wpe says so).
    """

    def __init__(self, taps=10, delay=2, iterations=3, psd_context=0, statistics_mode="full"):
        self.taps, self.delay, self.iterations = taps, delay, iterations
        self.psd_context, self.statistics_mode = psd_context, statistics_mode

    def _run(self, Y: torch.Tensor) -> torch.Tensor:
        if self.statistics_mode not in ("full", "valid"):
            raise ValueError(self.statistics_mode)
        D, T, F = Y.shape
        Y = Y.contiguous()
        lib = _lib.load()
        nbytes = lib.tssep_wpe_workspace_bytes(D, T, F, self.taps)
        if nbytes < 0:
            _lib.check(-1, "tssep_wpe_workspace_bytes")
        ws = torch.empty((nbytes + 255,), dtype=torch.uint8, device=Y.device)
        ws = ws[(-ws.data_ptr()) % 256:][:nbytes]
        X = torch.empty_like(Y)
        torch_ops.op.wpe(Y, D, T, F, int(self.taps), int(self.delay), int(self.iterations), int(self.psd_context),
                         int(self.statistics_mode == "valid"), X, ws, nbytes)
        return X

    def __call__(self, Observation, inplace=False):
        if isinstance(Observation, np.ndarray):
            if not torch.cuda.is_available():
                raise RuntimeError("tssep_b200 needs a CUDA device (no CPU fallback)")
            out = self._run(torch.as_tensor(Observation).to(device="cuda", dtype=torch.complex64)).cpu().numpy()
            out = out.astype(Observation.dtype if np.iscomplexobj(Observation) else np.complex128)
            if inplace:
                Observation[...] = out
                return Observation
            return out
        if isinstance(Observation, torch.Tensor):
            _lib.require_cuda(Observation)
            return self._run(Observation.to(torch.complex64)).to(
                Observation.dtype if Observation.is_complex() else torch.complex64)
        raise NotImplementedError(type(Observation), Observation)


class ChannelWiseWPE(WPE):
    """Every channel dereverberated on its own (tssep/train/enhancer.py:348-367): the (D, T, F) signal is handed to
    ``WPE`` as one channel with D * F independent frequencies."""

    def __call__(self, Observation, inplace=False):
        D, T, F = Observation.shape
        if isinstance(Observation, np.ndarray):
            flat = np.ascontiguousarray(np.transpose(Observation, (1, 0, 2)).reshape(1, T, D * F))
            out = np.transpose(super().__call__(flat).reshape(T, D, F), (1, 0, 2))
            if inplace:
                Observation[...] = out
                return Observation
            return np.ascontiguousarray(out)
        flat = Observation.permute(1, 0, 2).reshape(1, T, D * F)
        return super().__call__(flat).reshape(T, D, F).permute(1, 0, 2).contiguous()


def _intervals(ai):
    """(start, end) frame pairs of one speaker's activity: an ``ArrayInterval``-like object (``normalized_intervals``),
    an iterable of pairs, or a boolean activity vector."""
    if hasattr(ai, "normalized_intervals"):
        return [(int(s), int(e)) for s, e in ai.normalized_intervals]
    a = np.asarray(ai.cpu() if isinstance(ai, torch.Tensor) else ai)
    if a.dtype == bool:
        d = np.diff(np.concatenate([[0], a.astype(np.int8), [0]]))
        return list(zip(np.flatnonzero(d == 1).tolist(), np.flatnonzero(d == -1).tolist()))
    return [(int(s), int(e)) for s, e in a.reshape(-1, 2)]


class ClassicBF_np(ABC):
    """Segment-wise mask-based beamformer of the evaluation (tssep/train/enhancer.py:370-590): for every speaker and
    every interval of its diarization, spatial covariance matrices of target / distortion over the frames of the
    interval (``_get_psd``, enhancer.py:267-289 -- including its symmetrisation ``(psd + psd^T) / 2``, which keeps the
    real part of the Hermitian matrix), Souden MVDR towards channel 0, applied to the interval.

    Same constructor and call contract as the reference; ``Observation`` (mics, T, F) and ``masks``
    (spk, 1, T, F) may be numpy (moved to the current CUDA device) or CUDA tensors; ``dia`` a list (one entry per
    speaker) of ``ArrayInterval``-like objects, (start, end) lists or boolean vectors.  The statistics, the D x D
    solves and the filtering run in the kernels of csrc/beamform.cu (f64 statistics and solve, f32 signal);
    ``pb_bss`` (absent third-party package) is restated for ``mvdr_souden``, ``ch0`` and ``ch1`` -- the eigenvector
    beamformers (``scaled_gev_atf+mvdr``, ``rank1_gev+mvdr_souden``, ``wmwf``) raise ``NotImplementedError``.
    """

    @classmethod
    def finalize_dogmatic_config(cls, config):
        # as the reference (enhancer.py:420-427); WPE is activated by a {'factory': WPE} entry for pre_wpe / segment_wpe
        config["distortion_mask"] = {"factory": "tssep_b200.enhancer_distortion_mask.SumCrossTalker"}

    def __init__(self, bf="mvdr_souden", masking=False, masking_eps=0, distortion_mask=None, pre_wpe: "WPE" = None,
                 segment_wpe: "WPE" = None, mask_power=1):
        super().__init__()
        if distortion_mask is None:   # the reference's finalize_dogmatic_config default
            from .enhancer_distortion_mask import SumCrossTalker

            distortion_mask = SumCrossTalker()
        self.bf, self.masking, self.masking_eps = bf, masking, masking_eps
        self.distortion_mask, self.mask_power = distortion_mask, mask_power
        self.pre_wpe, self.segment_wpe = pre_wpe, segment_wpe

    def _beamform(self, m2: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
        """m2 (2, T', F) target / distortion weights of one speaker, Y (D, T', F) -> (T', F)."""
        D, T, F = Y.shape
        if self.bf in ("ch0", "ch1"):
            enh = Y[int(self.bf[2])]
        else:
            m = m2 if self.mask_power == 1 else m2 ** self.mask_power
            m = m.float().contiguous()[None, None]                       # (Z=1, K=1, 2, T', F)
            Yc = Y.contiguous()[None]
            psd = torch.empty((1, 2, F, D * (D + 1) // 2, 2), dtype=torch.float64, device=Y.device)
            torch_ops.op.bf_psd(Yc, m, 1, 1, 2, D, T, F, psd)
            psd[..., 1] = 0.0                                            # (psd + psd^T) / 2 of a Hermitian matrix
            w = torch.empty((1, 1, F, D), dtype=torch.complex64, device=Y.device)
            torch_ops.op.bf_mvdr_souden(psd, 1, 1, 2, D, F, 0, float(torch.finfo(torch.float64).tiny), w)
            out = torch.empty((1, 1, T, F), dtype=torch.complex64, device=Y.device)
            torch_ops.op.bf_apply(Yc, w, None, 1, 1, 2, D, T, F, 0.0, out)
            enh = out[0, 0]
        if self.masking:
            enh = enh * torch.clamp(m2[0].to(torch.float32), min=self.masking_eps)
        return enh

    def __call__(self, masks, Observation, dia, segment_bf=True, numpy_out=False):
        if self.bf not in ("mvdr_souden", "ch0", "ch1"):
            raise NotImplementedError(self.bf)
        if not torch.cuda.is_available():
            raise RuntimeError("tssep_b200 needs a CUDA device (no CPU fallback)")
        obs_np = isinstance(Observation, np.ndarray)
        dev = masks.device if isinstance(masks, torch.Tensor) and masks.is_cuda else (
            Observation.device if isinstance(Observation, torch.Tensor) and Observation.is_cuda else torch.device("cuda"))
        out_dtype = Observation.dtype
        masks = torch.as_tensor(masks).to(dev)
        Y = torch.as_tensor(Observation).to(dev)
        mics = Y.shape[0]
        assert mics >= 6 or self.bf in ("ch0", "ch1"), Y.shape     # all channels loaded (enhancer.py:466-472)
        if self.pre_wpe:
            Y = self.pre_wpe(Y)
        Y = Y.to(torch.complex64)
        masks = masks.permute(1, 0, 2, 3)                                # spk mask time freq -> mask spk time freq
        _, K, T, F = masks.shape
        if masks.shape[0] == 1 or self.bf == "ch0":
            masks = self.distortion_mask(masks[:1])
        else:
            assert masks.shape[0] == 2, masks.shape
            raise NotImplementedError(masks.shape)
        if dia is None:
            assert segment_bf is False and self.segment_wpe is None and numpy_out is True, (segment_bf, numpy_out)
            dia = [None] * K
        assert isinstance(dia, (tuple, list)), ("Expect list of ArrayInterval", type(dia), dia)
        ret = []
        out = torch.zeros((K, T, F), dtype=torch.complex64, device=dev) if numpy_out else None
        for k, ai in enumerate(dia):
            ret_spk = {}
            if segment_bf:
                for s, e in _intervals(ai):
                    Yl = Y[:, s:e]
                    if self.segment_wpe:
                        Yl = self.segment_wpe(Yl.contiguous())
                    ret_spk[(s, e)] = self._beamform(masks[:, k, s:e], Yl)
                    if numpy_out:
                        out[k, s:e] = ret_spk[(s, e)]
            else:
                assert self.segment_wpe is None, self.segment_wpe
                if ai is not None:
                    raise NotImplementedError("ToDo")                    # as the reference (enhancer.py:585)
                out[k] = self._beamform(masks[:, k], Y)
            ret.append(ret_spk)
        if numpy_out:
            if obs_np:
                return out.cpu().numpy().astype(out_dtype if np.issubdtype(out_dtype, np.complexfloating) else np.complex128)
            return out.to(out_dtype) if out_dtype.is_complex else out
        if obs_np:
            return [{k: v.cpu().numpy().astype(out_dtype) for k, v in r.items()} for r in ret]
        return ret
