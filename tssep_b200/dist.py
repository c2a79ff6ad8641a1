"""Data-parallel sharding of meetings across the GPUs of one box.

The path shards only across meetings: one process per GPU,
weights replicated, meetings dealt longest-first to the least-loaded rank, no
data-path collective.  The only exchange is one gather of the (tiny) diarization
results at the end -- ``torch.distributed`` over NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def assign_meetings(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment; returns the meeting indices of every rank.

    Deterministic (ties broken by index) so every rank computes the same plan without talking.
    """
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world_size
    plan = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += int(lengths[i])
    return plan


class SegmentGather:
    """One fixed-size all-gather of the per-meeting segment tables per step, with no host synchronisation.

    Every rank knows the whole plan (``assign_meetings`` is deterministic), so the padded message size and the
    scatter index of the gathered rows are computed once on the host: a step packs ``segments (m, K, S, 2)`` and
    ``counts (m, K)`` into one int32 buffer, issues ``all_gather_into_tensor`` asynchronously (NCCL orders it after
    the work already enqueued on the current stream) and returns; ``result()`` makes the current stream wait for the
    collective and scatters the rows to meeting order.  Nothing calls ``.item()`` or blocks the host on the GPU.
    """

    def __init__(self, plan: Sequence[Sequence[int]], rank: int, K: int, S: int, device, group=None,
                 n_total: Optional[int] = None):
        self.plan, self.rank, self.K, self.S, self.group = [list(p) for p in plan], rank, K, S, group
        self.world = len(plan)
        self.n_total = sum(len(p) for p in plan) if n_total is None else n_total  # ids absent from the plan stay zero
        self.m_max = max(1, max(len(p) for p in plan))
        self.width = K * S * 2 + K
        self.device = torch.device(device)
        # row r * m_max + j of the gathered buffer belongs to meeting plan[r][j]
        src, dst = [], []
        for r, ids in enumerate(self.plan):
            for j, mid in enumerate(ids):
                src.append(r * self.m_max + j)
                dst.append(mid)
        self._src = torch.tensor(src, dtype=torch.long, device=self.device)
        self._dst = torch.tensor(dst, dtype=torch.long, device=self.device)
        self._send = torch.zeros((self.m_max, self.width), dtype=torch.int32, device=self.device)
        self._recv = torch.zeros((self.world * self.m_max, self.width), dtype=torch.int32, device=self.device)
        self._work = None

    def start(self, segments: torch.Tensor, counts: torch.Tensor):
        """segments (m_local, K, S, 2) int32, counts (m_local, K) int32 of this rank's meetings, in plan order."""
        m = segments.shape[0]
        assert m == len(self.plan[self.rank]), (m, len(self.plan[self.rank]))
        if m:
            self._send[:m, : self.K * self.S * 2] = segments.reshape(m, -1)
            self._send[:m, self.K * self.S * 2:] = counts.reshape(m, -1)
        if self.world == 1 or not dist.is_initialized():
            self._recv.copy_(self._send)
            self._work = None
        else:
            self._work = dist.all_gather_into_tensor(self._recv, self._send, group=self.group, async_op=True)
        return self

    def result(self):
        """(segments (n_total, K, S, 2), counts (n_total, K)) in meeting order, on every rank."""
        if self._work is not None:
            self._work.wait()  # stream-side wait for NCCL; gloo (CPU tests) completes here
            self._work = None
        rows = self._recv.index_select(0, self._src)
        out = torch.zeros((self.n_total, self.width), dtype=torch.int32, device=self.device)
        out.index_copy_(0, self._dst, rows)
        KS2 = self.K * self.S * 2
        return out[:, :KS2].reshape(self.n_total, self.K, self.S, 2), out[:, KS2:].reshape(self.n_total, self.K)


def gather_segments(local_ids: Sequence[int], segments: torch.Tensor, counts: torch.Tensor, n_total: int,
                    group=None, plan: Optional[Sequence[Sequence[int]]] = None):
    """All-gathers per-meeting segment tables (convenience wrapper around ``SegmentGather``).

    segments (m_local, K, S, 2) int32, counts (m_local, K) int32 for the meetings ``local_ids``.
    Returns (segments (n_total, K, S, 2), counts (n_total, K)) on every rank, indexed by meeting id.
    ``plan`` = the meeting ids of every rank (``assign_meetings``); without it the ranks first exchange their id
    lists (one extra small collective and a host read).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = segments.device
    K, S = segments.shape[1], segments.shape[2]
    if plan is None:
        if world == 1:
            plan = [list(local_ids)]
        else:
            gathered: List[Optional[List[int]]] = [None] * world
            dist.all_gather_object(gathered, list(map(int, local_ids)), group=group)
            plan = gathered
    assert list(plan[rank]) == list(local_ids), (plan[rank], local_ids)
    return SegmentGather(plan, rank, K, S, dev, group, n_total=n_total).start(segments, counts).result()


# Relative cost of one dependent step of the tensor-memory recurrence at 8 / 16 / 32 / 64 rows per cluster, two row
# tiles per CTA, the faster of one / two sub-batches (measured at U = 300 on B200, profiles/r2_rec_ts_microbench.txt);
# the same ratios steer the kernel's own choice (csrc/lstm_ts.cu::kTsShapes).
TS_STEP_COST: Dict[int, float] = {8: 0.85, 16: 1.0, 32: 1.54, 64: 3.06}


def plan_recurrence_waves(n_items: int, rows_per_item: int, capacity: Dict[int, int], max_items: Optional[int] = None,
                          cost: Optional[Dict[int, float]] = None, wave_overhead: float = 0.02) -> List[int]:
    """Splits ``n_items`` meetings into waves for the K-rows-per-meeting recurrent layers.

    A launch costs T dependent steps whatever its batch; a step of a launch with ``rows`` rows costs
    ``min over w of cost[w] * ceil(rows / capacity[w])`` (``capacity[w]`` = rows one wave of co-resident clusters of
    width ``w`` holds, ``ops.recurrence_ts_capacity``).  Returns the wave sizes (largest first) that minimise the
    summed cost; ``max_items`` bounds a wave (memory of its input projections), ``wave_overhead`` is the relative
    cost of one more wave (launches, GEMM tails).
    """
    cost = TS_STEP_COST if cost is None else cost
    max_items = n_items if max_items is None else max(1, min(max_items, n_items))

    def wave_cost(w):
        rows = w * rows_per_item
        return min(cost[c] * -(-rows // capacity[c]) for c in capacity if capacity[c] > 0) + wave_overhead

    best = [0.0] + [float("inf")] * n_items
    choice = [0] * (n_items + 1)
    for m in range(1, n_items + 1):
        for w in range(1, min(m, max_items) + 1):
            c = best[m - w] + wave_cost(w)
            if c < best[m] - 1e-12:
                best[m], choice[m] = c, w
    waves, m = [], n_items
    while m > 0:
        waves.append(choice[m])
        m -= choice[m]
    return sorted(waves, reverse=True)
