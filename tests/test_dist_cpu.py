"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: meeting assignment and the
end-of-path gather of segment tables."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_assign_meetings_is_balanced_and_deterministic():
    from tssep_b200.dist import assign_meetings

    lengths = [600, 30, 590, 45, 300, 310, 10, 5]
    plan = assign_meetings(lengths, 3)
    assert sorted(i for p in plan for i in p) == list(range(8))
    loads = [sum(lengths[i] for i in p) for p in plan]
    assert max(loads) - min(loads) <= 45
    assert plan == assign_meetings(lengths, 3)
    assert assign_meetings([600] * 64, 8) == [list(range(r, 64, 8)) for r in range(8)]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tssep_b200.dist import SegmentGather, assign_meetings, gather_segments

    lengths = [100, 90, 80, 70, 60]
    plan = assign_meetings(lengths, world)
    mine = plan[rank]
    K, S = 2, 4
    seg = torch.zeros((len(mine), K, S, 2), dtype=torch.int32)
    cnt = torch.zeros((len(mine), K), dtype=torch.int32)
    for j, m in enumerate(mine):
        seg[j, :, 0, 0] = m * 10
        seg[j, :, 0, 1] = m * 10 + 5
        cnt[j] = m + 1
    out_s, out_c = gather_segments(mine, seg, cnt, len(lengths))
    ok = all(int(out_s[m, 0, 0, 0]) == m * 10 and int(out_s[m, 1, 0, 1]) == m * 10 + 5 and int(out_c[m, 0]) == m + 1
             for m in range(len(lengths)))
    # the serving form: plan known up front, one fixed-size collective per step, started asynchronously, reused
    sg = SegmentGather(plan, rank, K, S, "cpu")
    for step in range(2):
        s2, c2 = sg.start(seg + step, cnt).result()
        ok = ok and torch.equal(s2, out_s + step * (out_c[:, :, None, None] > 0)) and torch.equal(c2, out_c)
    ret[rank] = bool(ok) and len(mine) in (2, 3)
    dist.destroy_process_group()


def test_gather_segments_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_gather_segments_single_process():
    from tssep_b200.dist import gather_segments

    seg = torch.arange(2 * 2 * 3 * 2, dtype=torch.int32).reshape(2, 2, 3, 2)
    cnt = torch.tensor([[1, 2], [3, 0]], dtype=torch.int32)
    out_s, out_c = gather_segments([2, 0], seg, cnt, 3)
    assert torch.equal(out_s[2], seg[0]) and torch.equal(out_s[0], seg[1]) and int(out_s[1].abs().sum()) == 0
    assert out_c.tolist() == [[3, 0], [0, 0], [1, 2]]


def test_plan_recurrence_waves():
    """Wave planner of the K-rows-per-meeting layers: one launch while the rows fit one wave of clusters, else the
    cheapest partition under the measured step costs."""
    from tssep_b200.dist import plan_recurrence_waves

    cap = {8: 104, 16: 208, 32: 416}
    cost = {8: 1.0, 16: 1.35, 32: 2.3}
    assert plan_recurrence_waves(8, 8, cap, cost=cost) == [8]            # 64 rows: one wave of 8-row clusters
    assert plan_recurrence_waves(52, 8, cap, cost=cost) == [52]          # 416 rows: one wave of 32-row clusters
    w = plan_recurrence_waves(64, 8, cap, cost=cost)
    assert sum(w) == 64 and max(w) <= 52 and len(w) == 2                 # 512 rows do not fit one wave
    cap64, cost64 = {**cap, 64: 832}, {**cost, 64: 3.18}
    assert plan_recurrence_waves(64, 8, cap64, cost=cost64) == [64]      # ... unless clusters take 64 rows each
    assert plan_recurrence_waves(64, 8, cap, max_items=32, cost=cost) == [32, 32] or \
        sum(plan_recurrence_waves(64, 8, cap, max_items=32, cost=cost)) == 64
    assert plan_recurrence_waves(1, 8, cap, cost=cost) == [1]
