"""ctypes binding of ``libtssep_b200.so`` (include/tssep_b200.h).

The shared library is the product: if it is missing the import of any
``tssep_b200`` operator fails loudly -- there is no CPU or PyTorch fallback.
PyTorch is used for device memory, streams and ``torch.distributed`` only.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

# TSSEP_DEBUG_KNOBS=1 binds the debug library (tuning knobs / cycle counters compiled in, tssep_b200/build.py)
_LIB_PATH = Path(__file__).resolve().parent / "_lib" / (
    "libtssep_b200_dbg.so" if os.environ.get("TSSEP_DEBUG_KNOBS") == "1" else "libtssep_b200.so")
_lib = None

c_i64, c_i32, c_f32, c_vp = C.c_int64, C.c_int32, C.c_float, C.c_void_p


class GemmDesc(C.Structure):
    """Mirror of ``tssep_gemm_desc``."""

    _fields_ = [
        ("A", c_vp), ("lda", c_i64), ("a_stride", c_i64), ("a_div", c_i32),
        ("B", c_vp), ("ldb", c_i64), ("b_stride", c_i64), ("b_mod", c_i32),
        ("bias", c_vp), ("bias_stride", c_i64),
        ("M", c_i64), ("N", c_i32), ("K", c_i32), ("batch", c_i32),
        ("alpha", c_f32), ("act", c_i32), ("mode", c_i32),
        ("out", c_vp), ("ldo", c_i64), ("out_stride", c_i64), ("out_div", c_i32), ("out_stride_hi", c_i64),
        ("mask", c_vp), ("plane_map", c_vp), ("n_blocks", c_i32), ("row_len", c_i32),
        ("impl", c_i32), ("max_ctas", c_i32),
    ]


EPI_F32, EPI_BF16, EPI_HEAD = 0, 1, 2
ABI_VERSION = 4  # TSSEP_ABI_VERSION of include/tssep_b200.h this binding was written against

_SIGNATURES = {
    "tssep_abi_version": ([], C.c_int),
    "tssep_device_info": ([C.POINTER(C.c_int)] * 3, C.c_int),
    "tssep_stft": ([c_vp, c_i64, c_i64, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp], C.c_int),
    "tssep_feature_stats": ([c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp],
                            C.c_int),
    "tssep_feature_write": ([c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32,
                             c_i32, c_vp, c_vp, c_i64, c_vp], C.c_int),
    "tssep_wpe_workspace_bytes": ([c_i32, c_i64, c_i32, c_i32], c_i64),
    "tssep_wpe": ([c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp], C.c_int),
    "tssep_pcm16": ([c_vp, c_i64, c_f32, c_vp, c_vp], C.c_int),
    "tssep_log1p_abs": ([c_vp, c_i64, c_vp, c_vp], C.c_int),
    "tssep_ipd": ([c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp], C.c_int),
    "tssep_cast_bf16": ([c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp], C.c_int),
    "tssep_instance_norm": ([c_vp, c_i64, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp], C.c_int),
    "tssep_fold_embedding": ([c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp],
                             C.c_int),
    "tssep_gemm": ([C.POINTER(GemmDesc), c_vp], C.c_int),
    "tssep_head_expand_t": ([c_vp, c_i64, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp], C.c_int),
    "tssep_blstm_recurrence": ([c_vp, c_i32, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp], C.c_int),
    "tssep_pack_whh": ([c_vp, c_vp, c_i32, c_i32, c_vp, c_vp], C.c_int),
    "tssep_blstm_recurrence_ts": ([c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp], C.c_int),
    "tssep_blstm_recurrence_ts_capacity": ([c_i32, c_i32, c_i32, c_i32], C.c_int),
    "tssep_pack_whh_ts": ([c_vp, c_vp, c_i32, c_i32, c_vp, c_vp], C.c_int),
    "tssep_blstm_recurrence_train": ([c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp], C.c_int),
    "tssep_blstm_recurrence_bwd": ([c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp], C.c_int),
    "tssep_pack_whh_bwd": ([c_vp, c_vp, c_i32, c_i32, c_vp, c_vp], C.c_int),
    "tssep_mask_istft": ([c_vp, c_i64, c_vp, c_i64, c_i64, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp,
                          c_i64, c_vp, c_vp], C.c_int),
    "tssep_bf_psd": ([c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i64, c_i32, c_vp, c_vp], C.c_int),
    "tssep_bf_mvdr_souden": ([c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, C.c_double, c_vp, c_vp], C.c_int),
    "tssep_bf_apply": ([c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_i64, c_i32, c_f32, c_vp, c_vp], C.c_int),
    "tssep_activity": ([c_vp, c_i64, c_i64, c_i32, c_i64, c_vp, c_vp], C.c_int),
    "tssep_median_threshold": ([c_vp, c_i64, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp], C.c_int),
    "tssep_segments": ([c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp], C.c_int),
    "tssep_stft_vad": ([c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i64, c_vp, c_vp], C.c_int),
}

EXPORTED_SYMBOLS = ["tssep_last_error", *_SIGNATURES]


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m tssep_b200.build` "
            "(tssep_b200 has no CPU / PyTorch fallback)."
        )
    lib = C.CDLL(os.fspath(_LIB_PATH))
    lib.tssep_last_error.restype = C.c_char_p
    lib.tssep_last_error.argtypes = []
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    got = lib.tssep_abi_version()
    if got != ABI_VERSION:
        # a stale .so left behind by an older checkout would be bound with mismatched argument layouts
        raise RuntimeError(f"{_LIB_PATH} reports ABI version {got}, this binding needs {ABI_VERSION}: rebuild it with "
                           "`python -m tssep_b200.build --force`")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().tssep_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({code}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


class StreamHandle(int):
    """``cudaStream_t`` as an integer that remembers its device: ``call`` makes that device current for the launch."""

    device_index: int = -1

    def __new__(cls, handle: int, device_index: int):
        obj = super().__new__(cls, handle)
        obj.device_index = device_index
        return obj


def stream_of(t: torch.Tensor = None) -> StreamHandle:
    """The current stream of the tensor's device (not of the current device: they differ when a model lives on
    cuda:1 while cuda:0 is current)."""
    dev = t.device if t is not None else torch.device("cuda", torch.cuda.current_device())
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return StreamHandle(torch.cuda.current_stream(idx).cuda_stream, idx)


def require_cuda(*tensors):
    """No CPU fallback, and every operand of one call lives on one device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "tssep_b200 operators run on CUDA tensors only (no CPU fallback); got a tensor on " f"{t.device}"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tssep_b200 operator called with tensors on different devices: {dev} and {t.device}")


# Optional per-call device timing (bench.py): when a list is installed here every C-ABI call is
# bracketed by CUDA events recorded on the current stream; `launch_count` counts kernel launches.
_timeline = None
launch_count = 0


def set_timeline(timeline):
    global _timeline
    _timeline = timeline


def call(name: str, *args, detail: str = None):
    """Calls one C-ABI entry point; ``detail`` only refines the timeline entry (e.g. the GEMM shape).

    The kernels launch on the CURRENT device (occupancy queries, function attributes and the SM count are read from
    it as well), so the device of the stream argument (a ``StreamHandle`` from ``stream_of``) is made current for
    the call when it is not already."""
    global launch_count
    fn = getattr(load(), name)
    launch_count += 1
    dev = next((a.device_index for a in args if isinstance(a, StreamHandle)), -1)
    if dev >= 0 and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            _call_on_current_device(fn, name, args, detail)
        return
    _call_on_current_device(fn, name, args, detail)


def _call_on_current_device(fn, name, args, detail):
    if _timeline is None:
        check(fn(*args), name)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    check(fn(*args), name)
    end.record()
    _timeline.append((name, start, end, detail))
