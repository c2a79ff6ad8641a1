"""GPU parity: RNNP_packed (input GEMM + cluster recurrence + projection) vs torch.nn.LSTM/Linear.

bf16 operands, fp32 accumulation and cell state.  Tolerance: the reference output is O(1);
SURVEY.md §8d measured ~2.5e-4 logit error for full bf16 operand rounding at random init, we
allow 1e-2 absolute on RNNP outputs (sequence-long accumulation of bf16 rounding of h)."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

pytestmark = pytest.mark.gpu


def _pair(idim, units, hdim, seed=0):
    from tssep_b200.rnnp import RNNP_packed

    torch.manual_seed(seed)
    ref = O._RNNP(idim, units, hdim).eval()
    mine = RNNP_packed(idim, 1, units, hdim, 0).eval()
    mine.load_state_dict(ref.state_dict(), strict=True)
    return ref, mine.cuda()


@pytest.mark.parametrize("idim,units,hdim,shape", [
    (33, 2, 3, (6, 33)),            # tests/test_exp.py reduced sizes
    (64, 10, 12, (50, 64)),         # golden config sizes
    (80, 40, 42, (3, 316, 80)),     # toy config, batch 3
    (96, 40, 42, (2, 5, 100, 96)),  # 4-D input (batch, speaker, time, feat)
    (128, 128, 64, (9, 200, 128)),  # cluster of 4, two batch tiles
    (160, 300, 320, (8, 300, 160)), # full-size units: cluster of 8
    (160, 300, 320, (1, 1000, 160)),
])
def test_rnnp_matches_torch(cuda, idim, units, hdim, shape):
    ref, mine = _pair(idim, units, hdim)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    with torch.no_grad():
        want = ref(x)
        got = mine(x.to(cuda)).cpu()
    assert got.shape == want.shape
    err = (got - want).abs().max().item()
    assert err < 1e-2, err


@pytest.mark.parametrize("idim,units,hdim,shape", [
    (64, 40, 42, (3, 100, 64)),      # one CTA, one k-atom, 3 of 32 rows used
    (96, 64, 48, (40, 150, 96)),     # exactly 64 units, two row groups
    (80, 128, 64, (33, 120, 80)),    # cluster of 2
    (160, 300, 320, (64, 200, 160)), # full size: cluster of 5, partial last k-atom
    (160, 300, 320, (5, 300, 160)),
])
def test_rnnp_tcgen05_recurrence_matches_torch(cuda, monkeypatch, idim, units, hdim, shape):
    """The shared-memory / tcgen05 recurrence (csrc/lstm_tc.cu), forced for every row count."""
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", "tc")
    ref, mine = _pair(idim, units, hdim)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = ref(x)
        got = mine(x.to(cuda)).cpu()
    err = (got - want).abs().max().item()
    assert err < 1e-2, err
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", "regs")
    with torch.no_grad():
        got2 = mine(x.to(cuda)).cpu()
    assert (got2 - want).abs().max().item() < 1e-2


@pytest.mark.parametrize("layout", ["bt", "rows"])
@pytest.mark.parametrize("rows_per_cluster", ["16", "32"])
@pytest.mark.parametrize("idim,units,hdim,shape", [
    (64, 40, 42, (3, 100, 64)),      # one CTA, 3 of 16 rows used
    (96, 64, 48, (40, 150, 96)),     # exactly 64 units, two row groups, second one partly computed
    (80, 128, 64, (33, 120, 80)),    # cluster of 2
    (160, 300, 320, (64, 200, 160)), # full size: cluster of 5, partial last k-atom
    (160, 300, 320, (5, 300, 160)),
    (72, 10, 12, (17, 60, 72)),      # one k-step, Up = 16
])
def test_rnnp_tmem_recurrence_matches_torch(cuda, monkeypatch, layout, rows_per_cluster, idim, units, hdim, shape):
    """The tensor-memory recurrence (csrc/lstm_ts.cu), forced for every row count, both cluster widths, both
    G / H layouts (GEMM "BT" tiles with rows ordered (group, t, b32); plain rows streamed by TMA)."""
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", "ts")
    monkeypatch.setenv("TSSEP_TS_ROWS", rows_per_cluster)
    monkeypatch.setenv("TSSEP_TS_LAYOUT", layout)
    ref, mine = _pair(idim, units, hdim)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = ref(x)
        got = mine(x.to(cuda)).cpu()
    err = (got - want).abs().max().item()
    assert err < 1e-2, err


def test_rnnp_stress_weights(cuda):
    """All weights x4 (saturating gates), as SURVEY.md §8d asks; looser bound, reported not hidden."""
    ref, mine = _pair(64, 40, 42)
    with torch.no_grad():
        for p in ref.parameters():
            p.mul_(4.0)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn((4, 400, 64), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = ref(x)
        got = mine(x.to(cuda)).cpu()
    err = (got - want).abs().max().item()
    print("stress max abs err", err, "ref max", want.abs().max().item())
    assert err < 0.15 * max(1.0, want.abs().max().item()), err


def test_rnnp_multilayer_and_repr(cuda):
    from tssep_b200.rnnp import RNNP_packed

    torch.manual_seed(0)
    mine = RNNP_packed(64, 3, 40, 42, 0).eval()
    assert "LSTM(64, 40, batch_first=True, bidirectional=True)" in repr(mine)
    # torch reference of the same stack
    mods = list(mine.net)
    x = torch.randn((2, 120, 64), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        h = x
        for m in mods:
            h = m(h)[0] if isinstance(m, torch.nn.LSTM) else m(h)
        got = mine.cuda()(x.to(cuda)).cpu()
    assert (got - h).abs().max().item() < 1e-2
