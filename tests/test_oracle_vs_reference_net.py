"""Pins the oracle to the REFERENCE'S OWN CODE for the branches its doctests pin by shape only.

Two legs:
* live: ``tssep/train/net.py`` + ``rnnp.py`` + ``feature_extractor_torchaudio.py`` imported from ``/root/reference``
  behind the stand-ins of ``tests/ref_stub.py`` (build container only; skipped where the tree is absent), same
  weights, same ``np.random`` state -> the oracle must equal the reference EXACTLY (``== 0.0``);
* golden: outputs of that same reference run, committed as ``tests/golden/reference_net_goldens.npz``
  (``scripts/make_reference_goldens.py``), compared wherever the tests run.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O
from tests import ref_stub as RS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_net_goldens.npz")
FIELDS = ("mask", "logit", "embedding", "vad_mask", "vad_logit")

needs_reference = pytest.mark.skipif(not RS.available(), reason="reference tree not present on this machine")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _oracle_from_state(name, state):
    orc = O.OracleMaskEstimator(**RS.CASES[name]).eval()
    res = orc.load_state_dict(state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return orc


@needs_reference
@pytest.mark.parametrize("name", list(RS.CASES))
@pytest.mark.parametrize("batched", [False, True])
def test_oracle_equals_reference_mask_estimator(name, batched):
    ns = RS.load()
    torch.manual_seed(0)
    ref = ns.net.MaskEstimator_v2(aux_net=None, **RS.CASES[name]).eval()
    orc = _oracle_from_state(name, ref.state_dict())
    xs, aux = RS.case_inputs(name, batched)
    np.random.seed(3)
    with torch.no_grad():
        want = ref(xs, RS.aux_argument(aux, batched))
    np.random.seed(3)
    with torch.no_grad():
        got = orc(xs, RS.aux_argument(aux, batched))
    for f in FIELDS:
        w, g = getattr(want, f), getattr(got, f)
        if w is None:
            assert g is None, f
            continue
        assert w.shape == g.shape, (f, w.shape, g.shape)
        assert (w - g).abs().max().item() == 0.0, (name, batched, f)


@pytest.mark.parametrize("name", list(RS.CASES))
@pytest.mark.parametrize("batched", [False, True])
def test_oracle_equals_committed_reference_goldens(golden, name, batched):
    prefix = f"{name}/state/"
    state = {k[len(prefix):]: torch.tensor(golden[k]) for k in golden.files if k.startswith(prefix)}
    orc = _oracle_from_state(name, state)
    xs, aux = RS.case_inputs(name, batched)
    np.random.seed(3)
    with torch.no_grad():
        got = orc(xs, RS.aux_argument(aux, batched))
    tag = "batched" if batched else "single"
    seen = 0
    for f in FIELDS:
        key = f"{name}/{tag}/{f}"
        g = getattr(got, f)
        if key not in golden.files:
            assert g is None, key
            continue
        seen += 1
        # same arithmetic, same library: exact on the machine that made the goldens, rounding-level elsewhere
        np.testing.assert_allclose(g.numpy(), golden[key], rtol=0, atol=2e-6, err_msg=key)
    assert seen >= 2


@needs_reference
def test_oracle_equals_reference_torch_mfcc(golden):
    ns = RS.load()
    fe = ns.mfcc.TorchMFCC(size=1024, shift=256, window="hann")
    X = torch.tensor(golden["mfcc/X"])
    tables = O.MFCCTables()
    for x in (X[0], X):  # 2-D, and 3-D with torchaudio's batch-coupled top_db
        want = fe.stft_to_feature(x)
        got = O.mfcc_feature(x, tables)
        assert (want - got).abs().max().item() == 0.0


def test_oracle_mfcc_equals_committed_goldens(golden):
    X = torch.tensor(golden["mfcc/X"])
    tables = O.MFCCTables()
    np.testing.assert_allclose(O.mfcc_feature(X[0], tables).numpy(), golden["mfcc/single"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(O.mfcc_feature(X, tables).numpy(), golden["mfcc/batched"], rtol=0, atol=1e-4)
    # the coupling is real: item 0 of the batched call differs from the single call where the cut-off bites
    assert np.abs(golden["mfcc/batched"][0] - golden["mfcc/single"]).max() > 1.0


def test_oracle_instance_norms_equal_committed_goldens(golden):
    x = torch.tensor(golden["norm/x"])
    for dim in (-1, -2, 0):
        np.testing.assert_allclose(O.instance_norm(x, dim=dim).numpy(), golden[f"norm/v1/dim{dim}"], atol=1e-6)
        np.testing.assert_allclose(O.instance_norm(x, dim=dim, unbiased=True).numpy(), golden[f"norm/v1u/dim{dim}"],
                                   atol=1e-6)
    for md, nd in ((-1, -1), (-2, -2), (-2, -1)):
        np.testing.assert_allclose(O.instance_norm_v2(x, md, nd).numpy(), golden[f"norm/v2/{md}/{nd}"], atol=1e-6)


@needs_reference
def test_oracle_instance_norms_equal_reference():
    ns = RS.load()
    x = torch.tensor(np.random.RandomState(1).randn(4, 9, 5).astype(np.float32))
    for dim in (-1, -2, 0):
        assert (ns.net.InstanceNorm(dim=dim)(x) - O.instance_norm(x, dim=dim)).abs().max().item() == 0.0
    for md, nd in ((-1, -1), (-2, -2), (-2, -1)):
        assert (ns.net.InstanceNorm_v2(md, nd)(x) - O.instance_norm_v2(x, md, nd)).abs().max().item() == 0.0
