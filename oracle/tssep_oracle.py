"""CPU oracle: a restatement of the reference TS-SEP inference path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference
(merlresearch/tssep) cannot be imported here because padertorch / paderbox /
lazy_dataset / sacred are absent, so this module restates the arithmetic with
the very library calls the reference itself makes (``torch.nn.LSTM``,
``torch.nn.Linear``, ``torchaudio`` mel/dB/DCT, ``torch.fft``) plus a
restatement of the padertorch/paderbox STFT (padertorch==0.0.1,
paderbox==0.0.8, pinned in the reference's requirements.txt:16-17).

Parity status
-------------
* PINNED by the reference's own doctest goldens (``tests/test_oracle_goldens.py``):
  STFT, Log1pMaxNormAbsSTFT, ``cat`` conditioning, random speaker permutation,
  RNNP stack, ``tf`` head, sigmoid, Masking, iSTFT, LogMAE, DummyReader.
* pinned by shape / parameter-count goldens only: TorchMFCC, ``mul``
  conditioning, ``ts_vad`` speaker-concat layer, ``num_averaged_permutations>1``,
  ``output_resolution='t'``, ``explicit_vad``.
* PARITY UNPINNED: diarization post-processing (threshold / median smoothing /
  segment extraction) does not exist in the reference; the spec in
  ``diarize_reference`` is ours.

Every function cites the reference file:line it follows (paths relative to
the reference repository root).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence

import numpy as np
import scipy.signal
import torch

# --------------------------------------------------------------------------
# STFT / iSTFT  (padertorch.contrib.cb.feature_extractor.STFT, paderbox
# transform.module_stft; call sites tssep/train/model.py:504 and :661-664)
# --------------------------------------------------------------------------


def analysis_window(name: str, window_length: int) -> np.ndarray:
    """Periodic (DFT-even) window: ``scipy.signal.windows.<name>(L + 1)[:-1]``."""
    fn = getattr(scipy.signal.windows, name)
    return fn(window_length + 1)[:-1]


def synthesis_window(name: str, window_length: int, shift: int) -> np.ndarray:
    """Biorthogonal synthesis window ``w[n] / sum_k w[(n mod R) + kR]**2``."""
    assert window_length % shift == 0, (window_length, shift)
    w = analysis_window(name, window_length)
    denom = (w.reshape(window_length // shift, shift) ** 2).sum(axis=0)
    return w / np.tile(denom, window_length // shift)


def num_frames(num_samples: int, window_length: int, shift: int, pad=True, fading=True) -> int:
    """paderbox ``_samples_to_stft_frames`` (used at tssep/util/utils.py:32-38)."""
    if fading:
        num_samples = num_samples + 2 * (window_length - shift)
    frames = (num_samples - window_length + shift) / shift
    return int(math.ceil(frames)) if pad else int(math.floor(frames))


def _frame_signal(x, window_length, shift, pad, fading, xp):
    if fading:
        p = window_length - shift
        if xp is np:
            x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(p, p)])
        else:
            x = torch.nn.functional.pad(x, (p, p))
    n = x.shape[-1]
    if pad:
        t = int(math.ceil((n - window_length + shift) / shift))
        need = (t - 1) * shift + window_length
        if need > n:
            if xp is np:
                x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, need - n)])
            else:
                x = torch.nn.functional.pad(x, (0, need - n))
    else:
        t = (n - window_length + shift) // shift
    return x, t


def stft(x, size=1024, shift=256, window_length=None, window="blackman", pad=True, fading=True):
    """``fe.stft(signal)``: (..., N) -> (..., T, size//2+1).

    numpy input follows the numpy path (dtype preserved, golden at
    tssep/train/feature_extractor.py:197-202); torch input gives complex64
    for float32 signals (tssep/train/model.py:483).
    """
    if window_length is None:
        window_length = size
    w = analysis_window(window, window_length)
    if isinstance(x, np.ndarray):
        xs, t = _frame_signal(x, window_length, shift, pad, fading, np)
        idx = np.arange(window_length)[None, :] + shift * np.arange(t)[:, None]
        frames = xs[..., idx] * w
        return np.fft.rfft(frames, n=size, axis=-1)
    xs, t = _frame_signal(x, window_length, shift, pad, fading, torch)
    frames = xs.unfold(-1, window_length, shift)[..., :t, :]
    frames = frames * torch.as_tensor(w, dtype=x.dtype, device=x.device)
    return torch.fft.rfft(frames, n=size, dim=-1)


def istft(X, size=1024, shift=256, window_length=None, window="blackman", fading=True, num_samples=None):
    """``fe.istft(X, num_samples=N)``: (..., T, F) -> (..., N)."""
    if window_length is None:
        window_length = size
    v = synthesis_window(window, window_length, shift)
    is_np = isinstance(X, np.ndarray)
    if is_np:
        X = torch.as_tensor(X)
    frames = torch.fft.irfft(X, n=size, dim=-1)[..., :window_length]
    frames = frames * torch.as_tensor(v, dtype=frames.dtype)
    t = frames.shape[-2]
    total = (t - 1) * shift + window_length
    lead = frames.shape[:-2]
    out = torch.zeros(*lead, total, dtype=frames.dtype)
    # overlap-add, hop by hop (window_length // shift interleaved slabs)
    ov = window_length // shift
    for k in range(ov):
        sl = frames[..., k::ov, :]
        n_k = sl.shape[-2]
        if n_k == 0:
            continue
        flat = sl.reshape(*lead, n_k * window_length)
        start = k * shift
        out[..., start : start + n_k * window_length] += flat
    if fading:
        p = window_length - shift
        out = out[..., p : total - p]
    if num_samples is not None:
        out = out[..., :num_samples]
    return out.numpy() if is_np else out


# --------------------------------------------------------------------------
# Features
# --------------------------------------------------------------------------


def log1p_maxnorm_feature(X, statistics_axis="tf"):
    """``Log1pMaxNormAbsSTFT.stft_to_feature`` (tssep/train/feature_extractor.py:233-263)."""
    if isinstance(X, np.ndarray):
        s = np.abs(X)
        axis = {"tf": (-2, -1), "t": -2, "f": -1}[statistics_axis]
        norm = np.amax(s, keepdims=True, axis=axis)
        return np.log1p(s * ((np.e - 1) / norm))
    s = X.abs()
    dim = {"tf": (-2, -1), "t": -2, "f": -1}[statistics_axis]
    norm = torch.amax(s, keepdim=True, dim=dim)
    s = s * ((np.e - 1) / norm)
    return torch.log1p(s)


def interchannel_phase_differences(signal: np.ndarray, second_channel=None, concatenate=False):
    """tssep/train/feature_extractor.py:13-80.  signal (..., channels, frames, features) complex (numpy).

    ``interchannel_phase_differences_op`` is padertorch's (absent package, padertorch==0.0.1): the unit phasor of
    ``a * conj(b)``; pinned by the reference's doctest values (tests/test_ipd_features.py).  The partner channels are
    drawn from the global NumPy RNG exactly as feature_extractor.py:58-66 does."""
    import itertools

    signal = np.asarray(signal)
    if second_channel is None:
        D = signal.shape[-3]
        assert D >= 2, (D, signal.shape)
        pairs = list(itertools.permutations(range(D), 2))
        np.random.shuffle(pairs)
        second_channel = np.array(sorted(dict(pairs).items()))[:, 1]
    z = signal * np.conj(signal[..., second_channel, :, :])
    z = z / np.abs(z)
    if concatenate:
        return np.concatenate([np.abs(signal), z.real, z.imag], axis=-1)
    return z.real, z.imag


def log1p_abs_ipd_feature(X: np.ndarray) -> np.ndarray:
    """``Log1pAbsIPDSTFT.stft_to_feature`` (feature_extractor.py:96-109)."""
    return np.concatenate([np.log1p(np.abs(X)), *interchannel_phase_differences(X)], axis=-1)


def log1p_maxnorm_ipd_feature(X: np.ndarray, statistics_axis="tf") -> np.ndarray:
    """``Log1pMaxNormAbsIPDSTFT.stft_to_feature`` (feature_extractor.py:277-287)."""
    base = log1p_maxnorm_feature(np.asarray(X), statistics_axis)
    return np.concatenate([base, *interchannel_phase_differences(X)], axis=-1)


def mvn_log1p_abs_feature(X: np.ndarray) -> np.ndarray:
    """``MVNLog1pAbsSTFT.stft_to_feature`` (feature_extractor.py:154-168): log1p|X| minus its mean over the frames."""
    f = np.log1p(np.abs(X))
    return f - np.mean(f, axis=-2, keepdims=True)


class MFCCTables:
    """Constant tables of ``TorchMFCC`` (tssep/train/feature_extractor_torchaudio.py:22-85)."""

    def __init__(self, size=1024, sample_rate=16000, n_mfcc=40, dct_norm="ortho", f_min=40.0,
                 f_max=-400.0, n_mels=40, mel_norm=None, mel_scale="htk"):
        import torchaudio

        if f_max and f_max < 0:
            f_max = sample_rate + f_max
        self.top_db = 80
        self.to_db = torchaudio.transforms.AmplitudeToDB("power", self.top_db)
        self.mel = torchaudio.transforms.MelScale(
            n_mels, sample_rate, f_min, f_max, size // 2 + 1, mel_norm, mel_scale
        )
        self.dct = torchaudio.functional.create_dct(n_mfcc, n_mels, dct_norm)


def mfcc_feature(X: torch.Tensor, tables: MFCCTables) -> torch.Tensor:
    """``TorchMFCC.stft_to_feature`` (tssep/train/feature_extractor_torchaudio.py:93-106)."""
    power = X.transpose(-1, -2).abs().to(torch.float32) ** 2
    mel = tables.mel(power)
    mel = tables.to_db(mel)
    return torch.matmul(mel.transpose(-1, -2), tables.dct)


def concat_feature(X: torch.Tensor, tables: MFCCTables) -> torch.Tensor:
    """``ConcaternatedSTFTFeatures.stft_to_feature`` with fe1=TorchMFCC, fe2=Log1pMaxNormAbsSTFT
    (tssep/train/feature_extractor.py:352-360; order [mfcc | log1p-spectrum])."""
    return torch.concat([mfcc_feature(X, tables), log1p_maxnorm_feature(X)], dim=-1)


def instance_norm(x, dim=-1, unbiased=False):
    """``InstanceNorm.forward`` (tssep/train/net.py:280-285)."""
    std, mean = torch.std_mean(x, dim=dim, unbiased=unbiased, keepdim=True)
    return (x - mean) / std


def instance_norm_v2(x, mean_dim=-1, norm_dim=-1):
    """``InstanceNorm_v2.forward`` (tssep/train/net.py:322-330)."""
    x = x - torch.mean(x, dim=mean_dim, keepdim=True)
    norm = torch.linalg.norm(x, dim=norm_dim, keepdim=True) / np.sqrt(x.shape[norm_dim])
    return x / norm


# --------------------------------------------------------------------------
# Network (tssep/train/rnnp.py, tssep/train/net.py)
# --------------------------------------------------------------------------


class _RNNP(torch.nn.Module):
    """Single-layer ``RNNP_packed`` (tssep/train/rnnp.py:78-109 with elayers=1): BLSTM -> Linear."""

    def __init__(self, idim, cdim, hdim):
        super().__init__()
        self.net = torch.nn.ModuleList(
            [
                torch.nn.LSTM(idim, cdim, num_layers=1, bidirectional=True, batch_first=True),
                torch.nn.Linear(2 * cdim, hdim),
            ]
        )

    def forward(self, x):
        lead = None
        if x.dim() == 4:  # rnnp.py:124-136
            lead = x.shape[:2]
            x = x.reshape(lead[0] * lead[1], *x.shape[2:])
        h, _ = self.net[0](x)
        h = self.net[1](h)
        if lead is not None:
            h = h.reshape(*lead, *h.shape[1:])
        return h


@dataclasses.dataclass
class OracleOutput:
    mask: torch.Tensor
    logit: Optional[torch.Tensor]
    embedding: torch.Tensor = None
    vad_mask: torch.Tensor = None
    vad_logit: torch.Tensor = None


class OracleMaskEstimator(torch.nn.Module):
    """``MaskEstimator_v2`` restated (tssep/train/net.py:501-986).

    Parameter names equal the reference's (``pre_net.net.0.weight_ih_l0`` ...,
    ``post_net.linear2.bias``) so ``state_dict``s are interchangeable, and the
    modules are created in the reference's order so ``torch.manual_seed``
    reproduces its random init.
    """

    def __init__(self, *, idim=80, odim=None, layers=3, units=300, projs=320, nmask=1,
                 aux_net_output_size=100, combination="cat", ts_vad=False, output_resolution="tf",
                 random_speaker_order=True, num_averaged_permutations=1, explicit_vad=False,
                 input_normalizer=None, aux_normalizer=None):
        super().__init__()
        odim = idim if odim is None else odim
        self.odim, self.nmask, self.layers = odim, nmask, layers
        self.combination, self.ts_vad = combination, ts_vad
        self.output_resolution = output_resolution
        self.random_speaker_order = random_speaker_order
        self.num_averaged_permutations = num_averaged_permutations
        self.explicit_vad = explicit_vad
        self.input_normalizer, self.aux_normalizer = input_normalizer, aux_normalizer
        if not ts_vad:
            assert num_averaged_permutations == 1
        self.pre_net = _RNNP(idim, units, odim)
        first = odim + aux_net_output_size if combination == "cat" else odim
        post = torch.nn.Module()
        ts_factor = 1
        for l in range(layers):
            if l == layers - 1 and ts_vad is not False:
                ts_factor = ts_vad
            setattr(post, f"birnn{l}", _RNNP((first if l == 0 else projs) * ts_factor, units, projs))
        if output_resolution == "tf":
            out_features = (odim + int(explicit_vad)) * nmask * ts_factor
        else:
            assert output_resolution == "t" and not explicit_vad
            out_features = nmask * ts_factor
        setattr(post, f"linear{layers - 1}", torch.nn.Linear(projs, out_features))
        self.post_net = post

    # -- net.py:603-668 (post_net Sequential) --------------------------------
    def _post(self, xs):
        L = self.layers
        for l in range(L):
            if l == L - 1 and self.ts_vad is not False:
                # '... spk time feature -> ... 1 time (spk feature)'   net.py:606-612
                xs = xs.movedim(-3, -2)  # ... time spk feature
                xs = xs.reshape(*xs.shape[:-2], xs.shape[-2] * xs.shape[-1]).unsqueeze(-3)
            xs = getattr(self.post_net, f"birnn{l}")(xs)
            if l < L - 1:
                xs = torch.tanh(xs)
        xs = getattr(self.post_net, f"linear{L - 1}")(xs)
        K = self.ts_vad
        if self.output_resolution == "tf":
            fh = self.odim + int(self.explicit_vad)
            if K is False:
                # '... spk time (mask freq) -> ... spk mask time freq'
                xs = xs.reshape(*xs.shape[:-1], self.nmask, fh).movedim(-2, -3)
            else:
                # '... 1 time (spk mask freq) -> ... spk mask time freq'
                xs = xs.squeeze(-3)
                xs = xs.reshape(*xs.shape[:-1], K, self.nmask, fh)  # ... time spk mask freq
                xs = xs.movedim(-4, -2)  # ... spk mask time freq
        else:
            if K is False:
                # '... spk time mask -> ... spk mask time freq' (repeat)
                xs = xs.movedim(-1, -2)[..., None].expand(*xs.shape[:-2], self.nmask, xs.shape[-2], self.odim)
            else:
                xs = xs.squeeze(-3)
                xs = xs.reshape(*xs.shape[:-1], K, self.nmask).movedim(-3, -1)  # ... spk mask time
                xs = xs[..., None].expand(*xs.shape, self.odim)
        return xs.contiguous()

    def forward(self, xs, aux):
        # net.py:809-856
        if xs.dim() == 2:
            batched = False
            if self.random_speaker_order:
                perm = np.random.permutation(len(aux))
                iperm = np.argsort(perm)
                aux = [aux[i] for i in perm]
            aux = torch.stack(list(aux), dim=0)
            spk = aux.shape[0]
        elif xs.dim() == 3:
            batched = True
            if self.random_speaker_order:
                perm = [np.random.permutation(len(aux[0])) for _ in range(len(aux))]
                iperm = np.argsort(perm, axis=-1)
                aux = [[a[i] for i in p] for a, p in zip(aux, perm)]
            aux = torch.stack(
                [torch.stack(list(a), dim=0) if isinstance(a, (tuple, list)) else a for a in aux], dim=0
            )
            if self.aux_normalizer is not None:
                aux = self.aux_normalizer(aux)
            spk = aux.shape[1]
        else:
            raise RuntimeError(xs.shape)
        if self.input_normalizer is not None:
            xs = self.input_normalizer(xs)
        xs = self.pre_net(xs)  # net.py:860
        aux = aux.unsqueeze(-2)  # net.py:862-865
        if self.combination == "mul":  # net.py:871-874
            xs = xs[..., None, :, :] * aux
        elif self.combination == "cat":  # net.py:879-894
            t = xs.shape[-2]
            xs = torch.concat(
                [
                    xs[..., None, :, :].expand(*xs.shape[:-2], spk, t, xs.shape[-1]),
                    aux.expand(*aux.shape[:-2], t, aux.shape[-1]),
                ],
                dim=-1,
            )
        else:
            raise NotImplementedError(self.combination)
        trials = self.num_averaged_permutations
        if trials > 1:  # net.py:900-924
            if not batched:
                xs = xs[None]
            speakers = xs.shape[-3]
            idx = ((np.arange(speakers)[:, None] + np.arange(speakers)[None, :]) % speakers)[:trials, :].ravel()
            xs = xs[:, idx]
            xs = xs.reshape(xs.shape[0] * trials, speakers, *xs.shape[2:])
        logit = self._post(xs)  # net.py:926
        if trials > 1:  # net.py:928-955
            b = logit.shape[0] // trials
            logit = logit.reshape(b, trials * speakers, *logit.shape[2:])
            revert = np.argsort(idx.ravel())
            logit = logit[:, revert]
            logit = logit.reshape(b, speakers, trials, *logit.shape[2:]).mean(dim=2)
            if not batched:
                logit = logit.squeeze(0)
        if self.random_speaker_order:  # net.py:957-967
            if logit.dim() == 4:
                logit = logit[iperm]
            else:
                logit = logit[np.arange(len(logit))[:, None], iperm]
        if self.explicit_vad:  # net.py:969-980
            mask = torch.sigmoid(logit)
            vad = mask[..., 0]
            return OracleOutput(mask=mask[..., 1:] * vad[..., None], logit=None, vad_mask=vad,
                                vad_logit=logit[..., 0], embedding=aux)
        return OracleOutput(mask=torch.sigmoid(logit), logit=logit, embedding=aux)


def masking(mask: torch.Tensor, observation_stft: torch.Tensor, reference_channel=0) -> torch.Tensor:
    """``Masking.__call__`` (tssep/train/enhancer.py:73-100)."""
    obs = observation_stft[..., reference_channel, :, :]
    return obs[..., None, :, :] * torch.squeeze(mask, dim=-3)


def torch_bf(masks: torch.Tensor, observation_stft: torch.Tensor, reference_channel=0, masking=False, masking_eps=0.0,
             eps=None) -> torch.Tensor:
    """``TorchBF('mvdr_souden').__call__`` (tssep/train/enhancer.py:226-283), complex128 as the reference demands."""
    Y = observation_stft.to(torch.complex128)
    if masks.shape[-3] == 2:
        psds = torch.einsum("...kmtf,...dtf,...Dtf->...mkfdD", masks.to(torch.complex128), Y, Y.conj())
        target_psd, interference_psd = psds[..., 0, :, :, :, :], psds[..., 1, :, :, :, :]
    elif masks.shape[-3] == 1:
        m = torch.squeeze(masks, dim=-3).to(torch.complex128)
        target_psd = torch.einsum("...ktf,...dtf,...Dtf->...kfdD", m, Y, Y.conj())
        interference_psd = torch.einsum("...ktf,...dtf,...Dtf->...kfdD", 1 - m, Y, Y.conj())
    else:
        raise ValueError(masks.shape)
    phi = torch.linalg.solve(interference_psd, target_psd)
    lambda_ = torch.diagonal(phi, dim1=-2, dim2=-1).sum(-1)[..., None, None]
    eps = torch.finfo(lambda_.real.dtype).tiny if eps is None else eps
    mat = phi / torch.clamp(lambda_.real, min=eps)
    beamformer = mat[..., reference_channel]
    enh = torch.einsum("...kfd,...dtf->...ktf", beamformer.conj(), Y)
    if masking:
        enh = enh * torch.clamp(masks[..., :, 0, :, :], min=masking_eps)
    return enh


def sum_cross_talker(masks: np.ndarray, eps=0.0001) -> np.ndarray:
    """``SumCrossTalker.__call__`` (tssep/train/enhancer_distortion_mask.py:41-55)."""
    assert masks.shape[0] == 1, masks.shape
    noise = np.stack([np.sum(np.delete(masks, k, axis=1), axis=1) for k in range(masks.shape[1])], axis=1)
    return np.concatenate([masks, np.maximum(noise, eps)], axis=0)


def one_minus(masks: np.ndarray) -> np.ndarray:
    """``OneMinus.__call__`` (tssep/train/enhancer_distortion_mask.py:17-21)."""
    assert masks.shape[0] == 1, masks.shape
    return np.concatenate([masks, np.maximum(1 - masks, 0)], axis=0)


def _get_psd(mask, observation, mask_power=1):
    """``_get_psd`` (tssep/train/enhancer.py:267-289): mask (..., t), observation (..., d, t).  The last line adds the
    plain transpose, not the conjugate one: what remains of the Hermitian matrix is its real part."""
    if mask_power != 1:
        mask = mask ** mask_power
    psd = np.einsum("...t,...dt,...Dt->...dD", mask, observation, observation.conj()) / observation.shape[-1]
    return (psd + np.swapaxes(psd, -2, -1)) / 2


def mvdr_souden_vector(target_psd, noise_psd, ref_channel=0, eps=None):
    """``pb_bss.extraction.beamformer.get_mvdr_vector_souden`` (pb_bss@99eb6c8, absent third-party package; the
    published Souden MVDR): w = (Phi_n^-1 Phi_t)[:, ref] / max(Re trace(Phi_n^-1 Phi_t), eps)."""
    phi = np.linalg.solve(noise_psd, target_psd)
    lam = np.trace(phi, axis1=-1, axis2=-2)[..., None, None]
    eps = np.finfo(lam.real.dtype).tiny if eps is None else eps
    return (phi / np.maximum(lam.real, eps))[..., ref_channel]


def classic_bf_np(masks, Observation, dia, bf="mvdr_souden", masking=False, masking_eps=0, eps=0.0001, mask_power=1,
                  segment_bf=True):
    """``ClassicBF_np.__call__(..., numpy_out=True)`` (tssep/train/enhancer.py:455-590) with the default
    ``SumCrossTalker`` distortion mask: masks (spk, 1, T, F), Observation (mics, T, F), dia = per speaker a list of
    (start, end) frame intervals (the ``normalized_intervals`` of the reference's ArrayInterval) -> (spk, T, F).

    PARITY UNPINNED for the beamforming vector: ``pb_bss`` is absent, the reference's golden values for this class
    (SDR of its toy example, enhancer.py:411-418) need pb_bss / mir_eval to be reproduced.  ``_get_psd`` and the
    distortion masks are the reference's own code restated, the latter pinned by its doctest values."""
    masks = np.asarray(masks, dtype=np.float64)
    Y = np.transpose(np.asarray(Observation).astype(np.complex128), (2, 0, 1))       # freq mic time
    m = np.transpose(masks, (1, 0, 3, 2))                                            # mask spk freq time
    _, K, F, T = m.shape
    m = sum_cross_talker(m[:1], eps)
    out = np.zeros((K, T, F), dtype=np.complex128)
    for k in range(K):
        ivs = dia[k] if segment_bf else [(0, T)]
        for s, e in ivs:
            Yl = Y[:, :, s:e]
            if bf in ("ch0", "ch1"):
                est = Yl[:, int(bf[2]), :]
            else:
                pt = _get_psd(m[0, k, :, s:e], Yl, mask_power)
                pn = _get_psd(m[1, k, :, s:e], Yl, mask_power)
                w = mvdr_souden_vector(pt, pn, 0)
                est = np.einsum("...a,...at->...t", w.conj(), Yl)
            est = est.T
            if masking:
                est = est * np.maximum(m[0, k, :, s:e].T, masking_eps)
            out[k, s:e] = est
    return out


def _wpe_window_mean(x: np.ndarray, ctx: int) -> np.ndarray:
    """nara_wpe's ``window_mean(x, (ctx, ctx))``: mean over frames [t - ctx, t + ctx], normalised by the number of
    frames that exist."""
    if ctx == 0:
        return x
    T = x.shape[-1]
    c = np.concatenate([np.zeros(x.shape[:-1] + (1,)), np.cumsum(x, axis=-1)], axis=-1)
    lo = np.maximum(np.arange(T) - ctx, 0)
    hi = np.minimum(np.arange(T) + ctx, T - 1) + 1
    return (c[..., hi] - c[..., lo]) / (hi - lo)


def wpe(Y: np.ndarray, taps=10, delay=2, iterations=3, psd_context=0, statistics_mode="full") -> np.ndarray:
    """``WPE.__call__`` (tssep/train/enhancer.py:292-345): Y (D, T, F) complex -> dereverberated (D, T, F).

    PARITY UNPINNED: the arithmetic lives in ``nara_wpe.wpe.wpe_v8`` (requirements.txt: nara_wpe>=0.0.11), a
    third-party package absent from the reference tree and from this image; the reference holds no golden values for
    it (its doctest only compares its own numpy and torch paths).  This restates nara_wpe's published algorithm
    (Drude et al., "NARA-WPE", ITG 2018; wpe_v6/wpe_v8): per frequency, ``iterations`` times,
    lambda_t = mean_d |X[d,t]|^2 (window mean over +-psd_context), floored at 1e-10 max_t lambda;
    R = sum_t Yt_t Yt_t^H / lambda_t, P = sum_t Yt_t Y_t^H / lambda_t over all frames ('full') or the frames
    >= delay + taps - 1 ('valid'); G = solve(R, P); X = Y - G^H Yt, where Yt stacks the frames delayed by
    delay ... delay + taps - 1 (zeros before the first frame).  float64 / complex128 throughout.
    """
    Y = np.asarray(Y).astype(np.complex128)
    D, T, F = Y.shape
    Yf = np.transpose(Y, (2, 0, 1))                                  # (F, D, T)
    Yt = np.zeros((F, taps, D, T), dtype=np.complex128)
    for tau in range(taps):
        sh = delay + tau
        if sh < T:
            Yt[:, tau, :, sh:] = Yf[:, :, :T - sh]
    Yt = Yt.reshape(F, taps * D, T)
    s0 = delay + taps - 1 if statistics_mode == "valid" else 0
    assert statistics_mode in ("full", "valid"), statistics_mode
    X = Yf.copy()
    for _ in range(iterations):
        power = _wpe_window_mean(np.mean(X.real ** 2 + X.imag ** 2, axis=-2), int(psd_context))   # (F, T)
        eps = 1e-10 * np.max(power, axis=-1, keepdims=True)
        inv = 1.0 / np.maximum(power, eps)
        Yw = Yt * inv[:, None, :]
        R = Yw[..., s0:] @ np.conj(np.swapaxes(Yt[..., s0:], -1, -2))
        P = Yw[..., s0:] @ np.conj(np.swapaxes(Yf[..., s0:], -1, -2))
        G = np.linalg.solve(R, P)
        X = Yf - np.conj(np.swapaxes(G, -1, -2)) @ Yt
    return np.transpose(X, (1, 2, 0))


def channel_wise_wpe(Y: np.ndarray, **kw) -> np.ndarray:
    """``ChannelWiseWPE.__call__`` (tssep/train/enhancer.py:348-367): every channel dereverberated on its own."""
    return np.stack([wpe(y[None], **kw)[0] for y in np.asarray(Y)])


def log_mae(estimate: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``LogMAE.loss_fn`` (tssep/train/loss.py:244-247)."""
    return torch.log10((estimate - target).abs().mean(dim=-1).sum(dim=-1))


# --------------------------------------------------------------------------
# Fake data backend (tssep/data.py)
# --------------------------------------------------------------------------


def staircase_vad(num_samples: int, num_speakers: int) -> np.ndarray:
    """``DummyReader._get_vad`` (tssep/data.py:34-56)."""
    vad = np.zeros((num_speakers, num_samples), dtype=bool)
    start = 0
    for i in range(num_speakers):
        end = num_samples * (i + 2) // (num_speakers + 1)
        vad[i, start:end] = True
        start = end - (end - start) // 2
    return vad


def dummy_example(seed: int, sample_rate=16000, aux_size=100, num_samples=None, num_speakers=8,
                  dataset="validate"):
    """``DummyReader.__call__.get_example`` (tssep/data.py:75-139), ``num_samples`` generalised
    (the reference fixes ``sample_rate * 5``)."""
    if num_samples is None:
        num_samples = sample_rate * 5
    rng = np.random.RandomState(seed)
    frequency = rng.randint(100, 7000, size=(3, num_speakers))
    time = np.arange(num_samples) / sample_rate
    early = np.empty((num_speakers, num_samples), dtype=np.float32)
    for k in range(num_speakers):  # same sum order over the 3 sinusoids as .sum(axis=0)
        acc = np.sin(2 * np.pi * frequency[0, k] * time)
        acc = acc + np.sin(2 * np.pi * frequency[1, k] * time)
        acc = acc + np.sin(2 * np.pi * frequency[2, k] * time)
        early[k] = acc.astype(np.float32)
    vad = staircase_vad(num_samples, num_speakers)
    early *= vad
    noise = 1 * rng.rand(1, num_samples).astype(np.float32)
    observation = early[:, None, :].sum(axis=0) + noise
    aux = np.zeros((num_speakers, aux_size), dtype=np.float32)
    scale = 7000 + 1
    for spk, fs in enumerate(frequency.T):
        for f in fs:
            f = (f * aux_size) // scale
            aux[spk, f : f + 2] = 1
    return {
        "example_id": f"dummy_id_{seed}",
        "num_samples": num_samples,
        "observation": observation,  # (1, N)
        "speaker_reverberation_early_ch0": early,  # (K, N)
        "vad": vad,
        "auxInput": aux,
        "dataset": dataset,
        "reference_channel": 0,
    }


# --------------------------------------------------------------------------
# End-to-end path (tssep/train/model.py:465-536, :661-664)
# --------------------------------------------------------------------------


@dataclasses.dataclass
class OracleForward:
    mask: torch.Tensor
    logit: torch.Tensor
    embedding: torch.Tensor
    stft_estimate: torch.Tensor
    time_estimate: torch.Tensor
    Observation: torch.Tensor
    Input: torch.Tensor


def forward_path(observation: torch.Tensor, aux, net: OracleMaskEstimator, *, feature="log1p",
                 tables: MFCCTables = None, size=1024, shift=256, window="hann") -> OracleForward:
    """``Model.forward`` + the iSTFT of ``Model.review``.

    ``observation``: (C, N) or batched (B, C, N) float32.  ``aux``: (K, A) tensor / list of K
    vectors, or for the batched case (B, K, A).
    """
    with torch.no_grad():
        X = stft(observation, size=size, shift=shift, window=window)
        Xr = X[..., 0, :, :]
        if feature == "log1p":
            inp = log1p_maxnorm_feature(Xr)
        elif feature == "concat":
            inp = concat_feature(Xr, tables)
        else:
            raise ValueError(feature)
        inp = inp.to(torch.float32)
        if observation.dim() == 2:
            aux_arg = [a for a in aux]
        else:
            aux_arg = [[a for a in item] for item in aux]
        out = net(inp, aux_arg)
        est = masking(out.mask, X, 0)
        time = istft(est, size=size, shift=shift, window=window, num_samples=observation.shape[-1])
    return OracleForward(out.mask, out.logit, out.embedding, est, time, X, inp)


# --------------------------------------------------------------------------
# Frame <-> sample index mapping and diarization post-processing
# (tssep/util/utils.py:11-129 are the only in-repo anchors; the mapping helpers
# live in paderbox which is absent, and thresholding / smoothing / segment
# extraction live in the external fgnt/tssep_data repo.  PARITY UNPINNED: the
# spec below is this project's own and is pinned only by our own tests.)
# --------------------------------------------------------------------------


def sample_to_frame_index(sample, window_length, shift, fading=True):
    """Frame whose window centre is nearest to ``sample`` (original-signal coordinates)."""
    p = (window_length - shift) if fading else 0
    return np.maximum(0, (np.asarray(sample) + p - window_length // 2 + shift // 2) // shift)


def frame_to_sample_index(frame, window_length, shift, fading=True):
    """Smallest sample index that ``sample_to_frame_index`` maps to ``frame``."""
    p = (window_length - shift) if fading else 0
    return np.maximum(0, np.asarray(frame) * shift - p + window_length // 2 - shift // 2)


def median_smooth(x: np.ndarray, width: int) -> np.ndarray:
    """Running median over the last axis, odd ``width``, edges replicated."""
    assert width % 2 == 1 and width >= 1
    if width == 1:
        return x.copy()
    h = width // 2
    xp = np.concatenate([np.repeat(x[..., :1], h, axis=-1), x, np.repeat(x[..., -1:], h, axis=-1)], axis=-1)
    win = np.lib.stride_tricks.sliding_window_view(xp, width, axis=-1)
    return np.sort(win, axis=-1)[..., h]


def diarize_reference(mask: np.ndarray, *, threshold=0.5, median_width=1, window_length=1024,
                      shift=256, fading=True, num_samples=None):
    """Our post-processing spec.

    mask (K, 1, T, F) -> activity (K, T) = mean over F of mask[:, 0];
    smooth = running median; active = smooth > threshold;
    segments = maximal runs [t0, t1) of active frames, reported as sample
    intervals [frame_to_sample(t0), min(frame_to_sample(t1), num_samples)).
    Returns (activity, smooth, active, segments) with segments a list (per
    speaker) of (start_sample, end_sample) int tuples.
    """
    act = mask[:, 0].astype(np.float32).mean(axis=-1, dtype=np.float32)
    sm = median_smooth(act, median_width)
    active = sm > np.float32(threshold)
    segments = []
    for k in range(active.shape[0]):
        a = np.concatenate([[False], active[k], [False]])
        d = np.diff(a.astype(np.int8))
        starts = np.nonzero(d == 1)[0]
        ends = np.nonzero(d == -1)[0]
        s0 = frame_to_sample_index(starts, window_length, shift, fading)
        s1 = frame_to_sample_index(ends, window_length, shift, fading)
        if num_samples is not None:
            s0 = np.minimum(s0, num_samples)
            s1 = np.minimum(s1, num_samples)
        segments.append([(int(a_), int(b_)) for a_, b_ in zip(s0, s1)])
    return act, sm, active, segments


# paderbox index helpers used by tssep/util/utils.py (paderbox==0.0.8 is absent: restated from SURVEY.md App. A,
# PARITY UNPINNED for these three functions; the control flow of utils.py around them is pinned by
# tests/test_vad_utils.py, which runs the reference's own utils.py on top of them)


def pb_samples_to_stft_frames(samples, size, shift, *, pad=True, fading=False):
    """``paderbox.transform.module_stft._samples_to_stft_frames``."""
    if fading:
        samples = samples + 2 * (size - shift)
    frames = (samples - size + shift) / shift
    return int(math.ceil(frames)) if pad else int(math.floor(frames))


def pb_sample_index_to_stft_frame_index(sample, window_length, shift, fading=True):
    """``paderbox.transform.module_stft.sample_index_to_stft_frame_index``: 0 below ``ceil(wl/2)``, then one frame
    per ``shift`` samples; ``+ ceil((wl - shift) / shift)`` frames of fading."""
    sample = np.asarray(sample)
    h = (window_length + 1) // 2
    frame = np.where(sample < h, 0, (sample - h) // shift + 1)
    if fading:
        frame = frame + int(math.ceil((window_length - shift) / shift))
    return frame


def pb_stft_frame_index_to_sample_index(frame, window_length, shift, fading=True, mode="first"):
    """Set inverse of ``pb_sample_index_to_stft_frame_index``: 'first' = smallest sample mapped to a frame >= ``frame``,
    'last' = largest sample mapped to ``frame`` (= first(frame + 1) - 1)."""
    frame = np.asarray(frame)
    pad = int(math.ceil((window_length - shift) / shift)) if fading else 0
    h = (window_length + 1) // 2

    def first(f):
        f = f - pad
        return np.where(f <= 0, 0, h + (f - 1) * shift)

    if mode == "first":
        return first(frame)
    if mode == "last":
        return np.maximum(first(frame + 1) - 1, 0)
    raise ValueError(mode)


def _runs(a):
    d = np.diff(np.concatenate([[False], np.asarray(a, dtype=bool), [False]]).astype(np.int8))
    return list(zip(np.nonzero(d == 1)[0].tolist(), np.nonzero(d == -1)[0].tolist()))


def stft_vad(vad: np.ndarray, window_length, shift, fading=True) -> np.ndarray:
    """Sample activity -> frame activity (tssep/util/utils.py:11-77)."""
    vad = np.asarray(vad, dtype=bool)
    t = pb_samples_to_stft_frames(vad.shape[-1], window_length, shift, pad=True, fading=fading)
    out = np.zeros(vad.shape[:-1] + (t,), dtype=bool)
    for idx in np.ndindex(vad.shape[:-1]):
        for s, e in _runs(vad[idx]):
            fs = int(pb_sample_index_to_stft_frame_index(s, window_length, shift, fading))
            fe = int(pb_sample_index_to_stft_frame_index(e, window_length, shift, fading))
            out[idx][fs:fe] = True
    return out


def istft_vad(vad: np.ndarray, window_length, shift, fading=True, num_samples=None):
    """Frame activity -> sample intervals (tssep/util/utils.py:80-129): nested list of merged ``(start, end)``."""
    vad = np.asarray(vad, dtype=bool)
    out = np.empty(vad.shape[:-1], dtype=object)
    for idx in np.ndindex(vad.shape[:-1]):
        iv = []
        for s, e in _runs(vad[idx]):
            a = int(pb_stft_frame_index_to_sample_index(s, window_length, shift, fading, mode="first"))
            b = int(pb_stft_frame_index_to_sample_index(e, window_length, shift, fading, mode="last"))
            if num_samples is not None:
                a, b = min(a, num_samples), min(b, num_samples)
            iv.append((a, b))
        out[idx] = iv
    return out.tolist()
