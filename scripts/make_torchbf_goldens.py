#!/usr/bin/env python
"""Writes tests/golden/reference_torchbf_goldens.npz with the outputs of the REFERENCE's own TorchBF
(tssep/train/enhancer.py, imported behind tests/ref_stub.py) on the seeded scenes of tests/test_beamformer.py.
Build container only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import ref_stub as RS  # noqa: E402
from tests.test_beamformer import CASES, scene  # noqa: E402

enh = RS.load_enhancer()
out = {}
for i, case in enumerate(CASES):
    masks, Y = scene(**case)
    ex = {"Observation": Y, "reference_channel": 0}
    out[f"{i}/enh"] = enh.TorchBF("mvdr_souden")(masks, ex, None).numpy()
    out[f"{i}/enh_masking"] = enh.TorchBF("mvdr_souden", masking=True, masking_eps=0.1)(masks, ex, None).numpy()
path = os.path.join(ROOT, "tests", "golden", "reference_torchbf_goldens.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
