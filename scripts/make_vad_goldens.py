#!/usr/bin/env python
"""Writes tests/golden/reference_vad_goldens.npz: stft_vad / istft_vad outputs on the seeded cases of
tests/test_vad_utils.py.  Run in the build container AFTER tests/test_vad_utils.py::test_reference_utils_agree_with_the_oracle
passes there (that test runs the reference's own tssep/util/utils.py and compares it with the oracle that produces
these arrays)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tssep_oracle as O  # noqa: E402
from tests.test_vad_utils import _cases  # noqa: E402

out = {}
for i, (wl, shift, fading, v) in enumerate(_cases()):
    frames = O.stft_vad(v, wl, shift, fading)
    iv = O.istft_vad(frames, wl, shift, fading)
    out[f"{i}/vad"], out[f"{i}/frames"] = v, frames
    out[f"{i}/intervals"] = np.array([[k, a, b] for k, row in enumerate(iv) for a, b in row], dtype=np.int64).reshape(-1, 3)
path = os.path.join(ROOT, "tests", "golden", "reference_vad_goldens.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
