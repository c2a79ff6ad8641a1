"""Drop-in STFT feature extractors backed by the sm_100a kernels.

Mirrors ``tssep/train/feature_extractor.py`` (``Log1pMaxNormAbsSTFT`` :183-263,
``ConcaternatedSTFTFeatures`` :290-367) and the ``STFT`` base class those star-import
from padertorch (``stft``, ``istft``, ``stft_to_feature``, ``__call__``, ``output_size``,
``frequencies``): same constructor arguments, attribute names and call contract.
All arithmetic runs in ``libtssep_b200.so``; inputs must be CUDA tensors (numpy
inputs are uploaded to the current CUDA device and the result is returned as numpy).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.signal
import torch

from . import _lib, torch_ops
from .configurable import Configurable

_TOP_DB = 80.0


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class STFT(Configurable):
    """``padertorch.contrib.cb.feature_extractor.STFT`` work-alike (paderbox STFT semantics:
    periodic window, ``fading`` pads ``window_length - shift`` zeros on both ends, ``pad``
    extends the tail to a whole frame)."""

    def __init__(self, size=1024, shift=256, window_length=None, pad=True, fading=True, output_size=None,
                 window="blackman"):
        if window_length is None:
            window_length = size
        self.size = size
        self.shift = shift
        self.window_length = window_length
        self.pad = pad
        self.fading = fading
        self.window = window
        self.output_size = self._get_output_size(output_size)
        self._tables = {}

    @classmethod
    def finalize_dogmatic_config(cls, config):
        if config["window_length"] is None:
            config["window_length"] = config["size"]
        if config["output_size"] is None:
            config["output_size"] = cls._default_output_size(config)

    @classmethod
    def _default_output_size(cls, config):
        return config["size"] // 2 + 1

    def _get_output_size(self, output_size):
        return self.frequencies if output_size is None else output_size

    @property
    def frequencies(self):
        return self.size // 2 + 1

    def __repr__(self):
        import inspect

        names = [p for p in inspect.signature(type(self)).parameters if not isinstance(getattr(self, p, None), STFT)]
        return f"{type(self).__name__}(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in names) + ")"

    # -- index helpers -------------------------------------------------------
    def num_frames(self, num_samples: int) -> int:
        n = num_samples + (2 * (self.window_length - self.shift) if self.fading else 0)
        return int(math.ceil((n - self.window_length + self.shift) / self.shift))

    def sample_index_to_frame_index(self, sample):
        p = (self.window_length - self.shift) if self.fading else 0
        return np.maximum(0, (np.asarray(sample) + p - self.window_length // 2 + self.shift // 2) // self.shift)

    def frame_index_to_sample_index(self, frame):
        p = (self.window_length - self.shift) if self.fading else 0
        return np.maximum(0, np.asarray(frame) * self.shift - p + self.window_length // 2 - self.shift // 2)

    # -- constant tables (derived caches, per device) --------------------------
    def _device_tables(self, device):
        key = (device.type, device.index)
        tab = self._tables.get(key)
        if tab is None:
            if self.window_length % self.shift != 0:
                raise ValueError("window_length must be a multiple of shift")
            w = getattr(scipy.signal.windows, self.window)(self.window_length + 1)[:-1]
            ov = self.window_length // self.shift
            denom = (w.reshape(ov, self.shift) ** 2).sum(axis=0)
            syn = w / np.tile(denom, ov)
            k = np.arange(self.size // 2)
            tw = np.exp(-2j * np.pi * k / self.size)
            tab = {
                "window": torch.tensor(w, dtype=torch.float32, device=device),
                "synwin": torch.tensor(syn, dtype=torch.float32, device=device),
                "twiddle": torch.tensor(np.stack([tw.real, tw.imag], -1), dtype=torch.float32, device=device),
            }
            self._tables[key] = tab
        return tab

    # -- transforms ------------------------------------------------------------
    @staticmethod
    def _to_cuda(x, dtype):
        if isinstance(x, np.ndarray):
            if not torch.cuda.is_available():
                raise RuntimeError("tssep_b200 needs a CUDA device (no CPU fallback)")
            return torch.as_tensor(x).to(device="cuda", dtype=dtype), True
        _lib.require_cuda(x)
        return x.to(dtype), False

    def stft(self, signal, _window="window"):
        """(..., N) float -> (..., T, F) complex64.  (``_window="synwin"``: frames weighted with the synthesis window, the
        adjoint of ``istft`` up to a per-bin scale -- used by ``tssep_b200.autograd.ISTFTFn``.)"""
        x, was_np = self._to_cuda(signal, torch.float32)
        if not self.pad:
            raise NotImplementedError("pad=False is not supported by the CUDA STFT")
        x = x.contiguous()
        n = x.shape[-1]
        lead = x.shape[:-1]
        t = self.num_frames(n)
        tab = self._device_tables(x.device)
        out = torch.empty((*lead, t, self.frequencies), dtype=torch.complex64, device=x.device)
        n_sig = int(np.prod(lead)) if lead else 1
        torch_ops.op.stft(x, n_sig, n, tab[_window], tab["twiddle"], self.size, self.shift, self.window_length,
                          int(bool(self.fading)), t, out)
        return out.cpu().numpy() if was_np else out

    def istft(self, X, num_samples=None, _fading=None):
        """(..., T, F) complex -> (..., N) float32.  (``_fading=False``: keep the leading / trailing
        ``window_length - shift`` samples, for a frame RANGE of a longer signal -- ``tssep_b200.eval``.)"""
        fading = self.fading if _fading is None else _fading
        Xc, was_np = self._to_cuda(X, torch.complex64)
        Xc = Xc.contiguous()
        lead = Xc.shape[:-2]
        t = Xc.shape[-2]
        assert Xc.shape[-1] == self.frequencies, (Xc.shape, self.frequencies)
        total = (t - 1) * self.shift + self.window_length
        if fading:
            total -= 2 * (self.window_length - self.shift)
        n = total if num_samples is None else min(int(num_samples), total)
        tab = self._device_tables(Xc.device)
        out = torch.empty((*lead, n), dtype=torch.float32, device=Xc.device)
        n_sig = int(np.prod(lead)) if lead else 1
        torch_ops.op.mask_istft(Xc, 0, None, 0, n_sig, 1, t, self.size, self.shift, self.window_length,
                                int(bool(fading)), tab["synwin"], tab["twiddle"], None, out, n, None)
        return out.cpu().numpy() if was_np else out

    # feature description consumed by `_compute_features`
    def _feature_parts(self):
        raise NotImplementedError(type(self))

    def stft_to_feature(self, stft_signals):
        X, was_np = self._to_cuda(stft_signals, torch.complex64)
        out = _compute_features(self, X, want_f32=True)["f32"]
        return out.cpu().numpy() if was_np else out

    def __call__(self, signal):
        return self.stft_to_feature(self.stft(signal))


def _compute_features(fe: STFT, X: torch.Tensor, want_f32=True, want_bf16=False, couple=None):
    """Runs the two feature passes for ``fe`` on X (..., T, F) complex64.

    couple=None follows torchaudio's AmplitudeToDB packing rule (the top_db cut-off is shared
    over dim -3 of a >2-D input, tssep/train/feature_extractor_torchaudio.py:66-71, :100);
    couple=False forces per-item statistics (independent meetings).
    Returns dict with 'f32' (..., T, Din) and/or 'bf16' (rows, ld) plus 'ld'.
    """
    parts = fe._feature_parts()
    mfcc, with_log1p = parts.get("mfcc"), parts.get("log1p", False)
    X = X.contiguous()
    lead = X.shape[:-2]
    t, f = X.shape[-2:]
    n_items = int(np.prod(lead)) if lead else 1
    dev = X.device
    n_mels = mfcc.n_mels if mfcc is not None else 0
    n_mfcc = mfcc.n_mfcc if mfcc is not None else 0
    din = n_mfcc + (f if with_log1p else 0)
    if couple is None:
        couple = mfcc is not None and len(lead) >= 1
    groups = [(0, n_items)]
    if couple and len(lead) >= 2:
        g = lead[-1]
        groups = [(i, g) for i in range(0, n_items, g)]
    keys = [torch.empty((n_items,), dtype=torch.int32, device=dev) for _ in range(2)]  # order-preserving max keys
    meldb = torch.empty((n_items, t, max(n_mels, 1)), dtype=torch.float32, device=dev) if n_mels else None
    mt = mfcc._mel_tables(dev) if mfcc is not None else None
    torch_ops.op.feature_stats(X, n_items, t * f, t, f, mt["mel_t"] if mt else None, mt["lo"] if mt else None,
                               mt["hi"] if mt else None, n_mels, keys[0], keys[1], meldb)
    out = {}
    f32 = torch.empty((*lead, t, din), dtype=torch.float32, device=dev) if want_f32 else None
    ld = _round_up(din, 64) if din >= 64 else _round_up(din, 8)  # ops.operand_ld
    bf16 = torch.empty((n_items * t, ld), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    xv = torch.view_as_real(X).reshape(n_items, t, f, 2)
    for start, count in groups:
        torch_ops.op.feature_write(xv[start:], count, t * f, t, f, keys[0][start:], keys[1][start:],
                                   meldb[start:] if meldb is not None else None, mt["dct"] if mt else None, n_mels, n_mfcc,
                                   int(with_log1p), _TOP_DB, int(bool(couple)),
                                   f32.reshape(n_items, t, din)[start:] if f32 is not None else None,
                                   bf16[start * t:] if bf16 is not None else None, ld)
    out["f32"], out["bf16"], out["ld"], out["din"] = f32, bf16, ld, din
    return out


class Log1pMaxNormAbsSTFT(STFT):
    """``log1p(|X| (e-1) / max_{t,f}|X|)`` in [0, 1] (tssep/train/feature_extractor.py:183-263)."""

    def __init__(self, size=1024, shift=256, window_length=None, pad=True, fading=True, output_size=None,
                 window="blackman", statistics_axis="tf"):
        super().__init__(size=size, shift=shift, window_length=window_length, pad=pad, fading=fading,
                         output_size=output_size, window=window)
        self.statistics_axis = statistics_axis

    def _feature_parts(self):
        if self.statistics_axis != "tf":
            raise NotImplementedError(
                f"statistics_axis={self.statistics_axis!r}: only 'tf' (the value of every shipped config, "
                "tssep/exp/init_cfg_common.yaml:44) is implemented on the CUDA path"
            )
        return {"log1p": True}


def interchannel_phase_differences(signal, second_channel=None, concatenate=False):
    """cos / sin of ``angle(channel * second_channel.conj())`` (tssep/train/feature_extractor.py:13-80).

    signal (..., channels, frames, features) complex.  Without ``second_channel`` the partner channels are sampled
    exactly as the reference does -- ``np.random.shuffle`` of all ordered channel pairs, last pair per first channel
    wins -- so the global NumPy RNG is consumed identically."""
    import itertools

    X, was_np = STFT._to_cuda(signal, torch.complex64)
    if X.dim() < 3:
        raise IndexError(tuple(X.shape))
    D = X.shape[-3]
    if second_channel is None:
        assert D >= 2, (D, X.shape)
        pairs = list(itertools.permutations(range(D), 2))
        np.random.shuffle(pairs)
        second_channel = np.array(sorted(dict(pairs).items()))[:, 1]
    X = X.contiguous()
    TF = X.shape[-2] * X.shape[-1]
    lead = X.numel() // (D * TF)
    sc = torch.as_tensor(np.asarray(second_channel, dtype=np.int32), device=X.device)
    cos = torch.empty(X.shape, dtype=torch.float32, device=X.device)
    sin = torch.empty_like(cos)
    torch_ops.op.ipd(X, lead, D, TF, sc, cos, sin)
    if concatenate:
        out = torch.cat([X.abs(), cos, sin], dim=-1)
        return out.cpu().numpy() if was_np else out
    return (cos.cpu().numpy(), sin.cpu().numpy()) if was_np else (cos, sin)


class Log1pAbsSTFT(STFT):
    """``log1p(|X|)`` (padertorch ``Log1pAbsSTFT``, the base of the two classes below)."""

    def stft_to_feature(self, stft_signals):
        X, was_np = self._to_cuda(stft_signals, torch.complex64)
        X = X.contiguous()
        out = torch.empty(X.shape, dtype=torch.float32, device=X.device)
        torch_ops.op.log1p_abs(X, X.numel(), out)
        return out.cpu().numpy() if was_np else out


class MVNLog1pAbsSTFT(Log1pAbsSTFT):
    """``log1p(|X|)`` minus its mean over the frames (tssep/train/feature_extractor.py:112-168): the utterance-level
    mean normalisation of the front end."""

    def __init__(self, size=1024, shift=256, window_length=None, pad=True, fading=True, output_size=None,
                 window="blackman", norm_means: bool = True, norm_vars: bool = False, eps: float = 1.0e-20):
        super().__init__(size=size, shift=shift, window_length=window_length, pad=pad, fading=fading,
                         output_size=output_size, window=window)
        self.norm_means, self.norm_vars, self.eps = norm_means, norm_vars, eps

    def stft_to_feature(self, stft_signals):
        if not self.norm_means or self.norm_vars:
            raise NotImplementedError()  # as the reference (feature_extractor.py:160-166)
        X, was_np = self._to_cuda(stft_signals, torch.complex64)
        feature = Log1pAbsSTFT.stft_to_feature(self, X)
        from . import ops

        out = ops.instance_norm(feature, dim=-2, mode=1)
        return out.cpu().numpy() if was_np else out


class Log1pAbsIPDSTFT(Log1pAbsSTFT):
    """``[log1p|X| , cos IPD , sin IPD]`` per channel (tssep/train/feature_extractor.py:83-109); input (channels, T, F)."""

    def _get_output_size(self, output_size):
        return (self.size // 2 + 1) * 3 if output_size is None else output_size

    def stft_to_feature(self, stft_signals):
        was_np = isinstance(stft_signals, np.ndarray)
        X, _ = self._to_cuda(stft_signals, torch.complex64)
        base = Log1pAbsSTFT.stft_to_feature(self, X)
        cos, sin = interchannel_phase_differences(X, concatenate=False)
        out = torch.cat([base, cos, sin], dim=-1)
        return out.cpu().numpy() if was_np else out


class NoFeatureSTFT(STFT):
    """Zero features (tssep/train/feature_extractor.py:171-180)."""

    def stft_to_feature(self, stft_signals):
        return stft_signals[..., :0]

    def _get_output_size(self, output_size):
        if output_size is None:
            return 0
        assert output_size == 0, (output_size, self.frequencies)
        return output_size


class ConcaternatedSTFTFeatures(STFT, torch.nn.Module):
    """Concatenation ``[fe1 | fe2]`` of two STFT features (tssep/train/feature_extractor.py:290-367).

    The fused CUDA path supports the shipped combination fe1=TorchMFCC, fe2=Log1pMaxNormAbsSTFT
    (tssep/exp/init_cfg_common.yaml:12-50) plus either of them alone.
    """

    @classmethod
    def finalize_dogmatic_config(cls, config):
        for fe in ["fe1", "fe2"]:
            for k in ["size", "shift", "pad", "fading", "window"]:
                config[fe][k] = config[k]
            if config["window_length"] is not None:
                config[fe]["window_length"] = config["window_length"]
        if config["window_length"] is None:
            config["window_length"] = config["size"]
        for fe in ["fe1", "fe2"]:
            config[fe]["window_length"] = config["window_length"]
        if config["output_size"] is None:
            config["output_size"] = config["fe1"]["output_size"] + config["fe2"]["output_size"]

    def __init__(self, fe1, fe2, output_size=None, size=1024, shift=256, window="blackman", window_length=None,
                 pad=True, fading=True):
        torch.nn.Module.__init__(self)
        self._pair = (fe1, fe2)
        STFT.__init__(self, size=size, shift=shift, window_length=window_length, pad=pad, fading=fading,
                      output_size=output_size, window=window)
        self.fe1 = fe1
        self.fe2 = fe2

    def _get_output_size(self, output_size):
        fe1, fe2 = self._pair
        if output_size is None:
            return fe1._get_output_size(None) + fe2._get_output_size(None)
        return output_size

    def _feature_parts(self):
        p1, p2 = self.fe1._feature_parts(), self.fe2._feature_parts()
        if "mfcc" in p1 and p2.get("log1p") and "mfcc" not in p2 and not p1.get("log1p"):
            return {"mfcc": p1["mfcc"], "log1p": True}
        raise NotImplementedError(
            "ConcaternatedSTFTFeatures: the CUDA path implements fe1=TorchMFCC, fe2=Log1pMaxNormAbsSTFT "
            f"(got {type(self.fe1).__name__}, {type(self.fe2).__name__})"
        )


class Log1pMaxNormAbsIPDSTFT(Log1pMaxNormAbsSTFT):
    """``[log1p-max-norm |X| , cos IPD , sin IPD]`` per channel (tssep/train/feature_extractor.py:266-287)."""

    def _get_output_size(self, output_size):
        if output_size is None:
            return (self.size // 2 + 1) * 3
        assert output_size == self.frequencies * 3, (output_size, self.frequencies * 3)
        return output_size

    def stft_to_feature(self, stft_signals):
        was_np = isinstance(stft_signals, np.ndarray)
        X, _ = self._to_cuda(stft_signals, torch.complex64)
        base = _compute_features(self, X, want_f32=True)["f32"]
        cos, sin = interchannel_phase_differences(X, concatenate=False)
        out = torch.cat([base, cos, sin], dim=-1)
        return out.cpu().numpy() if was_np else out
