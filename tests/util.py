"""Shared helpers for the parity tests."""
import numpy as np
import torch

from oracle import tssep_oracle as O


def make_pair(oracle_kwargs, product_cls_kwargs=None, seed=0, device="cuda"):
    """Oracle mask estimator and product MaskEstimator_v2 with identical weights."""
    from tssep_b200.net import MaskEstimator_v2

    torch.manual_seed(seed)
    ref = O.OracleMaskEstimator(**oracle_kwargs).eval()
    kw = dict(oracle_kwargs)
    kw.update(product_cls_kwargs or {})
    me = MaskEstimator_v2.new(kw).eval()
    missing = me.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, me.to(device)


def sdr_db(est: np.ndarray, tgt: np.ndarray) -> float:
    num = (tgt ** 2).sum()
    den = ((est - tgt) ** 2).sum() + 1e-20
    return float(10 * np.log10(num / den + 1e-20))
