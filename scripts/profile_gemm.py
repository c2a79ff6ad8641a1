#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM on the shapes of the C3 model (per meeting group)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meetings", type=int, default=2)
    ap.add_argument("--frames", type=int, default=37503)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", type=str, default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    Bm, T, K8, Up, P, F = a.meetings, a.frames, 8, 304, 320, 513
    shapes = [
        # name, batch, M, N, K, mode, a_div
        ("pre_in   bf16", 1, Bm * T, 8 * Up, 553, ops.EPI_BF16),
        ("pre_proj bf16", 1, Bm * T, F, 2 * Up, ops.EPI_BF16),
        ("b0_in    bf16", Bm * K8, T, 8 * Up, 513, ops.EPI_BF16),
        ("b0_proj  bf16", 1, Bm * K8 * T, P, 2 * Up, ops.EPI_BF16),
        ("b1_in    bf16", 1, Bm * K8 * T, 8 * Up, P, ops.EPI_BF16),
        ("b2_in    bf16", Bm * 2, T, 8 * Up, K8 * P, ops.EPI_BF16),
        ("head     head", Bm, T, K8 * 520, 2 * P, ops.EPI_HEAD),   # product: 513-float rows padded to 520 (net.py)
        # experiments on the output path of the head (not product shapes): dense 513-float rows, 512-float rows
        # (128-byte aligned), the same GEMM through the plain f32 epilogue, mask only
        ("head513  head", Bm, T, K8 * F, 2 * P, ops.EPI_HEAD),
        ("head512  head", Bm, T, K8 * 512, 2 * P, ops.EPI_HEAD),
        ("headf32  f32 ", Bm, T, K8 * F, 2 * P, ops.EPI_F32),
        ("headmask head", Bm, T, K8 * F, 2 * P, ops.EPI_HEAD),
    ]
    for name, batch, M, N, K, mode in shapes:
        if not a.only and name.split()[0] in ("head513", "head512", "headf32", "headmask"):
            continue
        F = 512 if name.startswith("head512") else (520 if name.startswith("head ") else 513)
        if a.only and a.only not in name:
            continue
        ld = ops.operand_ld(K)
        A = (torch.randn((batch * M if mode != ops.EPI_F32 or batch == 1 else batch * M, ld), device=dev) * 0.1).to(torch.bfloat16)
        B = (torch.randn((N, ld), device=dev) * 0.1).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        kw = dict(batch=batch, a_stride=M * ld, b_stride=0, b_mod=1, bias=bias)
        if mode == ops.EPI_HEAD:
            logit = torch.empty((batch * K8, M, F), device=dev)
            mask = torch.empty_like(logit)
            pm = torch.arange(batch * K8, dtype=torch.int32, device=dev)
            lo = None if name.startswith("headmask") else logit
            run = lambda: ops.gemm(A, ld, B, ld, M, N, K, lo, mode=mode, mask=mask, plane_map=pm, n_blocks=K8, row_len=F, **kw)
            out_bytes = (1 if lo is None else 2) * logit.numel() * 4
        else:
            ldo = ops.round_up(N, 8)
            out = torch.empty((batch * M, ldo), dtype=torch.float32 if mode == ops.EPI_F32 else torch.bfloat16, device=dev)
            run = lambda: ops.gemm(A, ld, B, ld, M, N, K, out, mode=mode, ldo=ldo, out_stride=M * ldo, **kw)
            out_bytes = out.numel() * out.element_size()
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
        fl = 2.0 * batch * M * N * K
        print(f"{name} batch={batch:3d} M={M:8d} N={N:5d} K={K:5d}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  "
              f"out {out_bytes / ms / 1e6:7.0f} GB/s  in {(A.numel() * 2) / ms / 1e6:6.0f} GB/s", flush=True)
        del A, B


if __name__ == "__main__":
    main()
