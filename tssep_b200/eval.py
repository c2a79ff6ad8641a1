"""Evaluation driver around the inference path.

The reference only holds the hooks of this stage -- ``Model.prepare_eval_dataset`` / ``collate_fn``
(tssep/train/model.py:339-452), the pre-computed ``Observation`` branch of ``Model.forward`` (model.py:498-502) and
``reader.data_hooks.pre_net`` (tssep/data.py:148-152); the driver itself lives in the external fgnt/tssep_data
repository.  What it does there, and here: read meetings, run TS-SEP on whole meetings, turn the masks into
diarization segments, re-run the enhancer segment by segment (a beamformer's statistics belong to one segment), and
write one audio file per segment plus an RTTM file.

Everything heavy runs through ``Model.separate_waves`` / the enhancer kernels; this module is host orchestration:
batching meetings of equal length (longest first, the reference's ``sort=True`` idea, model.py:198-219), sharding them
over ranks (``tssep_b200.dist``), and the writers.
"""
from __future__ import annotations

import dataclasses
import os
import wave
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .dist import assign_meetings
from .enhancer import Masking


@dataclasses.dataclass
class Segment:
    meeting_id: str
    speaker: int
    start: int               # samples, inclusive
    end: int                 # samples, exclusive
    audio: Optional[np.ndarray] = None   # float32 (end - start,)
    path: Optional[str] = None


def collate_fn(exs: Sequence[dict]) -> dict:
    """``Model.collate_fn`` (tssep/train/model.py:339-373): list of examples -> dict of stacked arrays; the reference
    channel must agree inside a batch and becomes a scalar."""
    keys = exs[0].keys()
    ex = {k: [e[k] for e in exs] for k in keys}
    for k in ("observation", "Input", "auxInput", "vad", "speaker_reverberation_early_ch0"):
        if k in ex:
            arr = np.array([np.asarray(v.cpu()) if torch.is_tensor(v) else np.asarray(v) for v in ex[k]])
            if arr.dtype != object:
                ex[k] = arr
    if "reference_channel" in ex:
        assert len(set(ex["reference_channel"])) == 1, ex["reference_channel"]
        ex["reference_channel"] = ex["reference_channel"][0]
    return ex


def prepare_eval_dataset(examples: Iterable[dict], batch_size: Optional[int] = None, sort: bool = True,
                         rank: int = 0, world_size: int = 1) -> List[List[dict]]:
    """Batches of equal-length meetings for this rank, longest first (an OOM shows up on the first batch).

    Mirrors what ``Model.prepare_dataset`` does with ``lazy_dataset`` (model.py:181-337): ``prepare`` (observation /
    targets out of ``audio_data``, ``reference_channel = 0``), optional sort by length, batching."""
    prepared = []
    for e in examples:
        r = dict(e)
        if "audio_data" in e:
            r["observation"] = e["audio_data"]["observation"]
            for k, v in e["audio_data"].items():
                r.setdefault(k, v)
        r.setdefault("reference_channel", 0)
        prepared.append(r)
    lengths = [int(np.shape(e["observation"])[-1]) for e in prepared]
    mine = assign_meetings(lengths, world_size)[rank]
    order = sorted(mine, key=lambda i: (-lengths[i], i)) if sort else list(mine)
    batches: List[List[dict]] = []
    for i in order:
        if batches and lengths[i] == int(np.shape(batches[-1][0]["observation"])[-1]) and (
                batch_size is None or len(batches[-1]) < batch_size):
            batches[-1].append(prepared[i])
        else:
            batches.append([prepared[i]])
    return batches


def rttm_lines(segments: Sequence[Segment], sample_rate: int) -> List[str]:
    """NIST RTTM: ``SPEAKER <file> 1 <onset> <duration> <NA> <NA> <speaker> <NA> <NA>``."""
    out = []
    for s in sorted(segments, key=lambda s: (s.meeting_id, s.start, s.speaker)):
        onset, dur = s.start / sample_rate, (s.end - s.start) / sample_rate
        out.append(f"SPEAKER {s.meeting_id} 1 {onset:.3f} {dur:.3f} <NA> <NA> spk{s.speaker} <NA> <NA>")
    return out


def write_wav(path: str, audio: np.ndarray, sample_rate: int, peak: Optional[float] = None):
    """16-bit PCM mono; ``peak`` (default: the segment's own maximum, at least 1) maps to full scale."""
    a = np.asarray(audio, dtype=np.float32)
    scale = max(1.0, float(np.abs(a).max()) if a.size else 1.0) if peak is None else peak
    pcm = np.clip(np.round(a / scale * 32767.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(pcm.tobytes())


class EvalDriver:
    """meetings in -> per-speaker segments (audio + RTTM) out.

    ``segment_enhancer``: enhancer applied per segment to the frames of the segment (``None``: the separated signal of
    the whole-meeting pass is cut -- exact for ``Masking``, whose output does not depend on the context).  With a
    beamformer (``tssep_b200.enhancer.TorchBF``) the spatial statistics are estimated on the segment (+ ``context``
    frames on both sides), which is the point of the segment-wise re-run.
    """

    def __init__(self, model, *, threshold: float = 0.5, median_width: int = 11, max_segments: int = 256,
                 min_segment_samples: int = 0, segment_enhancer=None, context: int = 0, out_dir: Optional[str] = None,
                 sample_rate: int = 16000, wave: Optional[int] = None, out_wave: Optional[int] = None):
        self.model, self.segment_enhancer, self.context = model, segment_enhancer, int(context)
        self.diar = dict(threshold=threshold, median_width=median_width, max_segments=max_segments)
        self.min_segment_samples, self.out_dir, self.sample_rate = min_segment_samples, out_dir, sample_rate
        self.wave, self.out_wave = wave, out_wave

    # -- one batch of equal-length meetings -------------------------------------------------------------------------
    @torch.no_grad()
    def process_batch(self, batch: Sequence[dict], device) -> List[Segment]:
        ex = collate_fn(batch)
        ref = ex.get("reference_channel", 0)
        obs_all = torch.as_tensor(ex["observation"], dtype=torch.float32).to(device)   # (M, C, N)
        aux = torch.as_tensor(ex["auxInput"], dtype=torch.float32).to(device)          # (M, K, A)
        ids = [str(e.get("example_id", i)) for i, e in enumerate(batch)]
        n = obs_all.shape[-1]
        fe = self.model.fe
        segments: List[Segment] = []
        X_all = None
        if self.segment_enhancer is not None:
            X_all = fe.stft(obs_all)                                                   # (M, C, T, F): all channels
        for lo, hi, out in self.model.separate_waves(obs_all[:, ref], aux, wave=self.wave, out_wave=self.out_wave,
                                                     diarize=self.diar, want_estimate=False, want_time=True):
            seg = out.segments.segments.cpu().numpy()
            cnt = out.segments.counts.cpu().numpy()
            for m in range(hi - lo):
                for k in range(seg.shape[1]):
                    for a, b in seg[m, k, :min(int(cnt[m, k]), seg.shape[2])]:
                        a, b = int(a), min(int(b), n)
                        if b - a <= self.min_segment_samples:
                            continue
                        if self.segment_enhancer is None:
                            audio = out.time_estimate[m, k, a:b]
                        else:
                            audio = self._enhance_segment(out.mask[m], X_all[lo + m], ref, k, a, b, n)
                        segments.append(Segment(ids[lo + m], k, a, b, audio.float().cpu().numpy()))
            del out
        return segments

    def _enhance_segment(self, mask, X, ref, k, a, b, n):
        """Re-runs ``segment_enhancer`` on the frames that cover samples [a, b) (+ context) and cuts the result."""
        fe = self.model.fe
        T = mask.shape[-2]
        f0 = max(0, int(fe.sample_index_to_frame_index(a)) - self.context)
        f1 = min(T, int(fe.sample_index_to_frame_index(max(a, b - 1))) + 1 + self.context)
        ex = {"Observation": X[:, f0:f1].contiguous(), "reference_channel": ref}
        est = self.segment_enhancer(mask[:, :, f0:f1].contiguous(), ex, self.model)        # (K, f1 - f0, F)
        # time axis of the cut: frame f0 starts at sample f0 * shift - pad of the unpadded signal
        pad = (fe.window_length - fe.shift) if fe.fading else 0
        start = f0 * fe.shift - pad
        y = fe.istft(est[k].to(torch.complex64), _fading=False)  # the range keeps its edge samples: they are not signal padding
        lo, hi = a - start, b - start
        out = torch.zeros(b - a, dtype=torch.float32, device=y.device)
        src_lo, src_hi = max(lo, 0), min(hi, y.shape[-1])
        if src_hi > src_lo:
            out[src_lo - lo:src_hi - lo] = y[src_lo:src_hi]
        return out

    # -- a whole evaluation set -----------------------------------------------------------------------------------------
    def run(self, examples: Iterable[dict], device="cuda", batch_size: Optional[int] = None, rank: int = 0,
            world_size: int = 1) -> Dict[str, List[Segment]]:
        """Processes the meetings of this rank; writes ``<out_dir>/audio/<meeting>_spk<k>_<start>_<end>.wav`` and
        ``<out_dir>/rank<r>.rttm`` when ``out_dir`` is set.  Returns {meeting id: segments}."""
        _lib.load()
        result: Dict[str, List[Segment]] = {}
        for batch in prepare_eval_dataset(examples, batch_size=batch_size, rank=rank, world_size=world_size):
            for s in self.process_batch(batch, device):
                result.setdefault(s.meeting_id, []).append(s)
        if self.out_dir is not None:
            os.makedirs(os.path.join(self.out_dir, "audio"), exist_ok=True)
            for sid, segs in result.items():
                for s in segs:
                    s.path = os.path.join(self.out_dir, "audio", f"{sid}_spk{s.speaker}_{s.start:09d}_{s.end:09d}.wav")
                    write_wav(s.path, s.audio, self.sample_rate)
            with open(os.path.join(self.out_dir, f"rank{rank}.rttm"), "w") as f:
                f.write("\n".join(rttm_lines([s for v in result.values() for s in v], self.sample_rate)) + "\n")
        return result
