"""Data-parallel sharding of meetings across the GPUs of one box.

The path shards only across meetings: one process per GPU,
weights replicated, meetings dealt longest-first to the least-loaded rank, no
data-path collective.  The only exchange is one gather of the (tiny) diarization
results at the end -- ``torch.distributed`` over NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

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


def gather_segments(local_ids: Sequence[int], segments: torch.Tensor, counts: torch.Tensor, n_total: int,
                    group=None):
    """All-gathers per-meeting segment tables.

    segments (m_local, K, S, 2) int32, counts (m_local, K) int32 for the meetings ``local_ids``.
    Returns (segments (n_total, K, S, 2), counts (n_total, K)) on every rank, indexed by meeting id.
    Ranks may own different numbers of meetings; tables are padded to the maximum before the
    collective.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = segments.device
    K, S = segments.shape[1], segments.shape[2]
    if world == 1:
        out_s = torch.zeros((n_total, K, S, 2), dtype=torch.int32, device=dev)
        out_c = torch.zeros((n_total, K), dtype=torch.int32, device=dev)
        idx = torch.as_tensor(list(local_ids), dtype=torch.long, device=dev)
        out_s[idx], out_c[idx] = segments, counts
        return out_s, out_c
    m_local = torch.tensor([len(local_ids)], dtype=torch.int64, device=dev)
    m_all = [torch.zeros_like(m_local) for _ in range(world)]
    dist.all_gather(m_all, m_local, group=group)
    m_max = int(max(int(m.item()) for m in m_all))
    pad_s = torch.zeros((m_max, K, S, 2), dtype=torch.int32, device=dev)
    pad_c = torch.zeros((m_max, K), dtype=torch.int32, device=dev)
    pad_i = torch.full((m_max,), -1, dtype=torch.int64, device=dev)
    if len(local_ids):
        pad_s[: len(local_ids)], pad_c[: len(local_ids)] = segments, counts
        pad_i[: len(local_ids)] = torch.as_tensor(list(local_ids), dtype=torch.int64, device=dev)
    all_s = [torch.empty_like(pad_s) for _ in range(world)]
    all_c = [torch.empty_like(pad_c) for _ in range(world)]
    all_i = [torch.empty_like(pad_i) for _ in range(world)]
    dist.all_gather(all_s, pad_s, group=group)
    dist.all_gather(all_c, pad_c, group=group)
    dist.all_gather(all_i, pad_i, group=group)
    out_s = torch.zeros((n_total, K, S, 2), dtype=torch.int32, device=dev)
    out_c = torch.zeros((n_total, K), dtype=torch.int32, device=dev)
    for s, c, i in zip(all_s, all_c, all_i):
        valid = i >= 0
        out_s[i[valid]], out_c[i[valid]] = s[valid], c[valid]
    return out_s, out_c
