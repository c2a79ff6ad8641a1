"""Writes tests/golden/oracle_toy_tssep.npz: a sample of the oracle's outputs on the toy TS-SEP
config (BASELINE configs[1]).  The reference itself cannot be imported (padertorch etc. absent),
so this fixture pins the *oracle* against drift; the oracle is pinned to the reference by the
doctest goldens in reference_doctest_goldens.json.  Run from the repo root:
    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tssep_oracle as O  # noqa: E402

torch.manual_seed(0)
net = O.OracleMaskEstimator(idim=553, odim=513, units=40, projs=42, combination="mul", ts_vad=8,
                            aux_net_output_size=513, num_averaged_permutations=2).eval()
e = O.dummy_example(0, aux_size=513)
np.random.seed(0)
out = O.forward_path(torch.tensor(e["observation"]), torch.tensor(e["auxInput"]), net, feature="concat",
                     tables=O.MFCCTables(), window="hann")
rng = np.random.RandomState(123)
index = np.stack([rng.randint(0, 8, 64), np.zeros(64, int), rng.randint(0, 316, 64), rng.randint(0, 513, 64)], -1)
np.savez(os.path.join(os.path.dirname(__file__), "oracle_toy_tssep.npz"), index=index,
         mask=out.mask.numpy()[tuple(index.T)], time=out.time_estimate.numpy()[:, ::4001],
         input=out.Input.numpy()[::37, ::29])
print("written")
