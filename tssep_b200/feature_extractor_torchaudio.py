"""Drop-in ``TorchMFCC`` (tssep/train/feature_extractor_torchaudio.py:11-106).

|X|^2 -> mel filterbank (htk, f_max = sample_rate - 400 by default) -> 10 log10
(clamp 1e-10) -> top_db=80 cut-off -> orthonormal DCT.  The constant tables are
built on the host with the same torchaudio functions the reference calls in its
constructor; the per-frame arithmetic runs in ``tssep_feature_stats`` /
``tssep_feature_write``.
"""
from __future__ import annotations

import numpy as np
import torch

from .feature_extractor import STFT


class _MelTable(torch.nn.Module):
    """Holds the filterbank under the reference's state_dict key ``mel_scale.fb``."""

    def __init__(self, fb):
        super().__init__()
        self.register_buffer("fb", fb)


class TorchMFCC(STFT, torch.nn.Module):
    @classmethod
    def _default_output_size(cls, config):
        return config["n_mfcc"]

    def __init__(self, size=400, shift=200, window_length=None, pad=True, fading=True, output_size=None,
                 window="hann", sample_rate: int = 16000, n_mfcc: int = 40, dct_norm: str = "ortho",
                 log_mels: bool = False, f_min: float = 40, f_max: float = -400, n_mels: int = 40,
                 mel_norm: str = None, mel_scale: str = "htk"):
        import torchaudio

        torch.nn.Module.__init__(self)
        self.n_mfcc = n_mfcc
        STFT.__init__(self, size=size, shift=shift, window_length=window_length, pad=pad, fading=fading,
                      output_size=output_size, window=window)
        self.sample_rate = sample_rate
        self.f_min = f_min
        if f_max and f_max < 0:
            f_max = sample_rate + f_max
        self.f_max = f_max
        self.n_mels = n_mels
        self.dct_norm = dct_norm
        self.mel_norm = mel_norm
        self.mel_scale_name = mel_scale
        self.top_db = 80  # fixed in torchaudio as well
        self.log_mels = log_mels
        if log_mels:
            raise NotImplementedError("log_mels=True is not implemented on the CUDA path (configs use False)")
        fb = torchaudio.functional.melscale_fbanks(size // 2 + 1, f_min, f_max, n_mels, sample_rate, mel_norm,
                                                   mel_scale)  # (F, n_mels)
        self.mel_scale = _MelTable(fb)
        dct_mat = torchaudio.functional.create_dct(n_mfcc, n_mels, dct_norm)  # (n_mels, n_mfcc)
        self.register_buffer("dct_mat", dct_mat)
        self._mel_cache = {}

    def _get_output_size(self, output_size):
        return self.n_mfcc if output_size is None else output_size

    def _mel_tables(self, device):
        key = (device.type, device.index)
        tab = self._mel_cache.get(key)
        if tab is None:
            fb = self.mel_scale.fb.detach().cpu().numpy()  # (F, n_mels)
            lo = np.zeros(self.n_mels, dtype=np.int32)
            hi = np.zeros(self.n_mels, dtype=np.int32)
            for m in range(self.n_mels):
                nz = np.nonzero(fb[:, m])[0]
                if len(nz):
                    lo[m], hi[m] = nz[0], nz[-1] + 1
            tab = {
                "mel_t": torch.tensor(np.ascontiguousarray(fb.T), dtype=torch.float32, device=device),
                "lo": torch.tensor(lo, device=device),
                "hi": torch.tensor(hi, device=device),
                "dct": self.dct_mat.detach().to(device=device, dtype=torch.float32).contiguous(),
            }
            self._mel_cache[key] = tab
        return tab

    def _feature_parts(self):
        return {"mfcc": self}
