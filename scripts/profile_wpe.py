#!/usr/bin/env python
"""Times WPE (tssep_wpe) and the segment-wise beamformer on LibriCSS-shaped input: 7 channels, T frames, 513 bins."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200.enhancer import WPE, ChannelWiseWPE, TorchBF  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=37503)
    ap.add_argument("--channels", type=int, default=7)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    D, T, F = a.channels, a.frames, 513
    Y = torch.complex(torch.randn((D, T, F), device=dev, generator=g), torch.randn((D, T, F), device=dev, generator=g))
    for t in range(3, 8):  # some late reverberation
        Y[:, t:] += 0.2 * Y[:, :-t].clone()
    audio_s = T * 256 / 16000
    for name, enh in (("WPE(taps=10, delay=2, iterations=3)", WPE()), ("WPE(iterations=1)", WPE(iterations=1)),
                      ("ChannelWiseWPE()", ChannelWiseWPE())):
        ms = timed(lambda: enh(Y), a.reps)
        DK = (1 if "Channel" in name else D) * enh.taps
        cmacs = enh.iterations * (D * F if "Channel" in name else F) * T * (DK * DK / 2 + DK * (1 if "Channel" in name else D)) * 4
        print(f"{name}: D={D} T={T} F={F}: {ms:8.2f} ms = {audio_s / (ms / 1e3):8.0f} audio-s/s, statistics {2 * cmacs / ms / 1e9:6.1f} TFLOP/s (f32 FMA)",
              flush=True)
    K = 8
    mask = torch.rand((K, 1, T, F), device=dev, generator=g)
    bf = TorchBF()
    ms = timed(lambda: bf(mask, {"Observation": Y, "reference_channel": 0}), a.reps)
    print(f"TorchBF(mvdr_souden): K={K} D={D} T={T}: {ms:8.2f} ms = {audio_s / (ms / 1e3):8.0f} audio-s/s", flush=True)


if __name__ == "__main__":
    main()
