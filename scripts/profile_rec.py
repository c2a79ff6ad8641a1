#!/usr/bin/env python
"""Micro-benchmark of the BLSTM recurrence kernels: microseconds per dependent step for a few batch sizes, cluster
widths and math modes, timed with CUDA events (also the ncu target).

With the debug library (TSSEP_DEBUG_KNOBS=1 at build AND run time) the per-phase cycle counters of two epilogue warps
of the tensor-memory kernel are printed too."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=300)
    ap.add_argument("--frames", type=int, default=4000)
    ap.add_argument("--rows", type=int, nargs="+", default=[1, 8, 16, 64, 128])
    ap.add_argument("--clusters", type=int, nargs="+", default=[0], help="rows per cluster (ts) / CTAs per cluster (regs)")
    ap.add_argument("--fast", type=int, nargs="+", default=[1])
    ap.add_argument("--ksplit", type=int, nargs="+", default=[0])
    ap.add_argument("--tiles", type=int, nargs="+", default=[2], help="row tiles per CTA of the tensor-memory kernel (0 = auto)")
    ap.add_argument("--subs", type=int, nargs="+", default=[0], help="sub-batches per cluster (0 = auto)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kernel", default="ts", choices=["regs", "ts"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    U = a.units
    Up = ops.round_up(U, 16)
    torch.manual_seed(0)
    w = (torch.rand((2, 4 * U, U), device=dev) - 0.5) * (2 / U ** 0.5)
    whh = ops.pack_whh(w[0], w[1], U, Up) if Up <= 320 else None
    wts = ops.pack_whh_ts(w[0].contiguous(), w[1].contiguous(), U, Up)
    debug = os.environ.get("TSSEP_DEBUG_KNOBS") == "1"

    def run(G, rows, C, fast, ks):
        if a.kernel == "ts":
            tiles, ksp, subs = ks
            return ops.blstm_recurrence_ts(G, wts, rows, a.frames, Up, fast_math=bool(fast), rows_per_cluster=C, k_split=ksp,
                                           tiles_per_cta=tiles, sub_batches=subs)
        return ops.blstm_recurrence(G, whh, rows, a.frames, Up, cluster=C, fast_math=bool(fast))

    for rows in a.rows:
        G = torch.empty((rows, a.frames, 8 * Up), device=dev, dtype=torch.bfloat16).normal_(0.0, 0.3)
        for C in a.clusters:
            for fast in a.fast:
                for ks in ([(tl, k, sb) for tl in a.tiles for k in a.ksplit for sb in a.subs
                            if not (tl == 1 and k == 1) and not (sb == 2 and (tl == 1 or k == 1 or C == 8))
                            and not (sb == 1 and C == 64)] if a.kernel == "ts" else [0]):
                    for _ in range(2):
                        run(G, rows, C, fast, ks)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.reps):
                        run(G, rows, C, fast, ks)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / a.reps / a.frames
                    ph = ""
                    if debug:
                        prof = torch.zeros(12, dtype=torch.int32, device=dev)
                        os.environ["TSSEP_REC_PROF"] = str(prof.data_ptr())
                        run(G, rows, C, fast, ks)
                        torch.cuda.synchronize()
                        del os.environ["TSSEP_REC_PROF"]
                        pc = prof.cpu().numpy().astype(float)
                        if a.kernel == "ts":
                            names = ["t0.wait", "t0.ld", "t0.math", "t0.send", "t1.wait", "t1.ld", "t1.math", "t1.send",
                                     "mma.pg", "mma.hwait", "mma.fence", "mma.issue"]
                            ph = " | cycles/step " + " ".join(f"{n}={v / a.frames:.0f}" for n, v in zip(names, pc))
                        else:
                            ph = " | cycles/step " + " ".join(
                                f"{n}={v / max(pc[5], 1):.0f}" for n, v in zip(["gwait", "hwait", "mma", "gates", "send"], pc[:5]))
                    print(f"U={U} rows={rows:4d} kernel={a.kernel} cluster={C} fast={fast} (tiles, ksplit, subs)={ks}: {us:.3f} us/step{ph}",
                          flush=True)
        del G


if __name__ == "__main__":
    main()
