#!/usr/bin/env python
"""Micro-benchmark / ncu target of the mask x STFT -> iSTFT kernel on LibriCSS-shaped meetings."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200.enhancer import Masking  # noqa: E402
from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meetings", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--speakers", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    fe = Log1pMaxNormAbsSTFT(size=1024, shift=256, window="hann")
    n = int(a.seconds * 16000)
    x = torch.randn((a.meetings, 1, n), device=dev)
    X = fe.stft(x)
    T, F = X.shape[-2:]
    mask = torch.rand((a.meetings, a.speakers, 1, T, 520), device=dev)[..., :F]   # rows at the head GEMM's pitch (net.py)
    act = torch.empty((a.meetings, a.speakers, T), device=dev)
    for label, act_out in (("without activity", None), ("with fused activity", act), ("without activity", None),
                           ("with fused activity", act)):
        run = lambda: Masking.apply(mask, X, 0, fe, want_estimate=True, want_time=True, num_samples=n, activity_out=act_out)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        nbytes = a.meetings * (8 * T * F + a.speakers * (4 * T * F + 8 * T * F + 4 * n))
        print(f"mask_istft {label}: meetings={a.meetings} speakers={a.speakers} T={T}: {ms:.3f} ms  "
              f"{nbytes / ms / 1e6:.0f} GB/s algorithmic ({nbytes / 1e9:.2f} GB)")


if __name__ == "__main__":
    main()
