#!/usr/bin/env python
"""Micro-benchmark / ncu target of the feature front end (STFT -> statistics -> feature rows) on LibriCSS-shaped
meetings: achieved GB/s of the three kernels against their algorithmic bytes."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200 import _lib  # noqa: E402
from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures, _compute_features  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meetings", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    fe = ConcaternatedSTFTFeatures.new({
        "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
        "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
        "size": 1024, "shift": 256, "window": "hann"})
    n = int(a.seconds * 16000)
    x = torch.rand((a.meetings, n), device=dev)

    def run():
        X = fe.stft(x)
        return _compute_features(fe, X, want_f32=True, want_bf16=True, couple=False)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    tl = []
    _lib.set_timeline(tl)
    for _ in range(a.reps):
        run()
    torch.cuda.synchronize()
    _lib.set_timeline(None)
    T, F, M = fe.num_frames(n), 513, a.meetings
    algo = {"tssep_stft": M * (4 * n + 8 * T * F), "tssep_feature_stats": M * (8 * T * F + 4 * T * 40),
            "tssep_feature_write": M * (8 * T * F + 4 * T * 40 + 6 * T * 553)}
    agg = {}
    for name, s, e, _ in tl:
        agg[name] = agg.get(name, 0.0) + s.elapsed_time(e) / a.reps
    for name, ms in agg.items():
        print(f"{name}: {ms:.3f} ms  {algo[name] / ms / 1e6:.0f} GB/s algorithmic ({algo[name] / 1e9:.2f} GB, {M} meetings)")


if __name__ == "__main__":
    main()
