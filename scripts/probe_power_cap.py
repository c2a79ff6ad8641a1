#!/usr/bin/env python
"""Does a latency-bound recurrence launch run slower right after a power-hungry GEMM burst (sw_power_cap clock dip),
and does a short idle gap before it help?  Prints us/step for: alone, after a burst, after burst + idle gap."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tssep_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
U, Up, rows, T = 300, 304, 104, 12000
torch.manual_seed(0)
w = (torch.rand((2, 4 * U, U), device=dev) - 0.5) * (2 / U ** 0.5)
wts = ops.pack_whh_ts(w[0].contiguous(), w[1].contiguous(), U, Up)
G = torch.empty((rows, T, 8 * Up), device=dev, dtype=torch.bfloat16).normal_(0.0, 0.3)
a = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)
b = torch.randn((8192, 8192), device=dev, dtype=torch.bfloat16)


def burst(ms):
    n = max(1, int(ms / 0.8))  # one 8192^3 bf16 matmul is about 0.8 ms
    for _ in range(n):
        torch.matmul(a, b)


def rec():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.blstm_recurrence_ts(G, wts, rows, T, Up)
    e1.record()
    return e0, e1


for _ in range(2):
    rec()
torch.cuda.synchronize()
time.sleep(0.5)
e = rec()
torch.cuda.synchronize()
print(f"alone:                 {e[0].elapsed_time(e[1]) * 1e3 / T:.3f} us/step")
for burst_ms in (10, 40, 150):
    for gap_ms in (0, 2, 10):
        time.sleep(0.5)
        burst(burst_ms)
        if gap_ms:
            torch.cuda._sleep(int(gap_ms * 1.9e6))  # cycles
        e = rec()
        torch.cuda.synchronize()
        print(f"burst {burst_ms:3d} ms, gap {gap_ms:2d} ms: {e[0].elapsed_time(e[1]) * 1e3 / T:.3f} us/step")
# sustained alternation, as in the bench
time.sleep(0.5)
ts = []
for i in range(6):
    burst(35)
    ts.append(rec())
torch.cuda.synchronize()
print("alternating 35 ms burst / recurrence:", " ".join(f"{x[0].elapsed_time(x[1]) * 1e3 / T:.3f}" for x in ts))
