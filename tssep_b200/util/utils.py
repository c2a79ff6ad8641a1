"""Drop-in ``stft_vad`` / ``istft_vad`` (tssep/util/utils.py:11-129): sample activity <-> STFT frame activity.

Same call signatures.  The reference walks the runs of every signal on the host and maps their bounds with paderbox's
``sample_index_to_stft_frame_index`` / ``stft_frame_index_to_sample_index`` (paderbox==0.0.8, absent in this image);
here both directions are one kernel each (``tssep_stft_vad``: a gather, ``tssep_segments`` with ``index_mode=1``: warp
ballots + prefix sums) on the index mapping restated in ``csrc/postproc.cu`` -- PARITY UNPINNED for the mapping itself,
the control flow of ``utils.py`` around it is pinned by ``tests/test_vad_utils.py``.

Differences from the reference, by necessity: ``istft_vad`` returns plain interval lists ``[(start, end), ...]`` (the
content of paderbox's ``ArrayInterval.normalized_intervals``) instead of ``ArrayInterval`` objects, and numpy inputs
make a round trip through the GPU (there is no CPU implementation).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib, torch_ops


def num_frames(num_samples: int, window_length: int, shift: int, fading: bool = True) -> int:
    """paderbox ``_samples_to_stft_frames(n, size=window_length, shift, pad=True, fading)`` (utils.py:31-38)."""
    n = num_samples + (2 * (window_length - shift) if fading else 0)
    return int(math.ceil((n - window_length + shift) / shift))


def _as_cuda_u8(x):
    if isinstance(x, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("tssep_b200 needs a CUDA device (no CPU fallback)")
        return torch.as_tensor(x.astype(np.uint8)).cuda(), "np"
    _lib.require_cuda(x)
    return (x != 0).to(torch.uint8), "torch"


def stft_vad(vad, window_length, shift, fading):
    """Moves a sample activity (..., N) to a frame activity (..., T)."""
    if isinstance(vad, (tuple, list)):
        return [stft_vad(v, window_length, shift, fading) for v in vad]
    if not isinstance(vad, (np.ndarray, torch.Tensor)):
        raise TypeError(vad)
    v, kind = _as_cuda_u8(vad)
    v = v.contiguous()
    n_samples = v.shape[-1]
    lead = v.shape[:-1]
    n = int(np.prod(lead)) if lead else 1
    T = num_frames(n_samples, window_length, shift, fading)
    out = torch.empty((*lead, T), dtype=torch.uint8, device=v.device)
    torch_ops.op.stft_vad(v, n, n_samples, window_length, shift, int(bool(fading)), T, out)
    if kind == "np":
        return out.cpu().numpy().astype(bool)
    return out.to(vad.dtype) if vad.dtype.is_floating_point else out.bool()  # the reference returns a float Tensor


def istft_vad(vad, window_length, shift, fading, num_samples=None, max_segments=4096):
    """Moves a frame activity (..., T) to sample intervals: an object array (numpy input) / nested list of
    ``[(start, end), ...]`` per signal, ``end`` exclusive.  ``num_samples`` clips the intervals."""
    if isinstance(vad, (tuple, list)):
        return [istft_vad(v, window_length, shift, fading, num_samples, max_segments) for v in vad]
    if not isinstance(vad, (np.ndarray, torch.Tensor)):
        raise TypeError(vad)
    v, _ = _as_cuda_u8(vad)
    v = v.contiguous()
    T = v.shape[-1]
    lead = v.shape[:-1]
    n = int(np.prod(lead)) if lead else 1
    seg = torch.zeros((n, max_segments, 2), dtype=torch.int32, device=v.device)
    cnt = torch.empty((n,), dtype=torch.int32, device=v.device)
    torch_ops.op.segments(v, n, T, window_length, shift, int(bool(fading)), -1 if num_samples is None else int(num_samples),
                          seg, cnt, max_segments, 1)
    seg_h, cnt_h = seg.cpu().numpy(), cnt.cpu().numpy()
    if int(cnt_h.max(initial=0)) > max_segments:
        raise ValueError(f"more than max_segments={max_segments} runs in one signal")
    flat = [[(int(a), int(b)) for a, b in s[:c]] for s, c in zip(seg_h, cnt_h)]
    data = np.empty(lead, dtype=object)
    for i, idx in enumerate(np.ndindex(*lead)):
        data[idx] = flat[i]
    return data.tolist() if lead else flat[0]
