#!/usr/bin/env python
"""Micro-benchmark of the BLSTM recurrence kernel: microseconds per recurrent step for a few
batch sizes / cluster sizes / math modes, timed with CUDA events (also the ncu target)."""
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
    ap.add_argument("--clusters", type=int, nargs="+", default=[0])
    ap.add_argument("--fast", type=int, nargs="+", default=[0, 1])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kernel", default="regs", choices=["regs", "tc", "ts", "ts_rows"])
    ap.add_argument("--g-bf16", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    U = a.units
    Up = ops.round_up(U, 16)
    torch.manual_seed(0)
    w = (torch.rand((2, 4 * U, U), device=dev) - 0.5) * (2 / U ** 0.5)
    whh = ops.pack_whh(w[0], w[1], U, Up)
    wimg = ops.pack_whh_tc(w[0].contiguous(), w[1].contiguous(), U, Up)
    wts = ops.pack_whh_ts(w[0].contiguous(), w[1].contiguous(), U, Up)

    def run(G, rows, C, fast):
        if a.kernel == "ts_rows":
            return ops.blstm_recurrence_ts(G, wts, rows, a.frames, Up, fast_math=bool(fast), rows_per_cluster=C, layout="rows")
        if a.kernel == "ts":
            return ops.blstm_recurrence_ts(G, wts, rows, a.frames, Up, fast_math=bool(fast), rows_per_cluster=C)
        if a.kernel == "tc":
            return ops.blstm_recurrence_tc(G, wimg, rows, a.frames, Up, fast_math=bool(fast))
        H = torch.empty((rows, a.frames, 2 * Up), dtype=torch.bfloat16, device=G.device)
        _lib.call("tssep_blstm_recurrence", G.data_ptr(), a.g_bf16, whh.data_ptr(), H.data_ptr(), rows, a.frames, Up, C, int(fast),
                  _lib.stream_of(G))
        return H

    for rows in a.rows:
        G = torch.empty((((rows + 31) // 32) * 32, a.frames, 8 * Up), device=dev,
                        dtype=torch.bfloat16 if a.g_bf16 else torch.float32).normal_(0.0, 0.3)
        for C in a.clusters:
            for fast in a.fast:
                for _ in range(2):
                    run(G, rows, C, fast)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    run(G, rows, C, fast)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / a.reps / a.frames
                prof = torch.zeros(8, dtype=torch.int32, device=dev)
                os.environ["TSSEP_REC_PROF"] = str(prof.data_ptr())
                run(G, rows, C, fast)
                torch.cuda.synchronize()
                del os.environ["TSSEP_REC_PROF"]
                pc = prof.cpu().numpy().astype(float)
                if a.kernel in ("ts", "ts_rows"):
                    names = ["t0.g", "t0.wait", "t0.math", "t0.send", "t1.g", "t1.wait", "t1.math", "t1.send"]
                    ph = " ".join(f"{n}={v / a.frames:.0f}" for n, v in zip(names, pc))
                elif a.kernel == "tc":
                    names = ["t0.wait", "t0.ld+act", "t0.cell", "t0.send", "t1.wait", "t1.ld+act", "t1.cell", "t1.send"]
                    ph = " ".join(f"{n}={v / a.frames:.0f}" for n, v in zip(names, pc))
                else:
                    ph = " ".join(f"{n}={v / max(pc[5], 1):.0f}" for n, v in zip(["gwait", "hwait", "mma", "gates", "send"], pc[:5]))
                print(f"U={U} rows={rows:4d} cluster={C} fast={fast}: {us:.3f} us/step | cycles/step {ph}", flush=True)
        del G


if __name__ == "__main__":
    main()
