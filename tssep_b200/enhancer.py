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
        m = masks.float().contiguous()
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
        torch_ops.op.mask_istft(obs, T * F, m, Z, K, T, fe.size, fe.shift, fe.window_length, int(bool(fe.fading)),
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
