"""Builds ``tssep_b200/_lib/libtssep_b200.so`` (the C-ABI of include/tssep_b200.h)
with plain nvcc for sm_100a.  No torch headers are involved, so the library has
a stable ABI and compiles in seconds; nvcc cross-compiles without a GPU.

    python -m tssep_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_lib"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
]
# TSSEP_DEBUG_KNOBS=1 selects the DEBUG library (libtssep_b200_dbg.so, built next to the product library): tuning /
# profiling knobs compiled in (environment variables read by the library, per-phase cycle counters of the
# recurrence).  The product library has none of them.  tssep_b200/_lib.py loads whichever this variable names.
DEBUG = os.environ.get("TSSEP_DEBUG_KNOBS") == "1"
if DEBUG:
    NVCC_FLAGS = NVCC_FLAGS + ["-DTSSEP_DEBUG_KNOBS"]
TAG = "_dbg" if DEBUG else ""
LIB = OUT_DIR / f"libtssep_b200{TAG}.so"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "tssep_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    stamp = OUT_DIR / f"build{TAG}.stamp"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src: Path) -> Path:
        obj = OUT_DIR / (src.stem + TAG + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(_sources()))) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
