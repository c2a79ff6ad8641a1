"""Thin typed wrappers around the operators of ``torch.ops.tssep_b200`` shared by several modules.

Nothing here computes on the host: each function validates, allocates the
output with torch (device memory only) and enqueues one kernel from
``libtssep_b200.so`` on the current CUDA stream through the custom-op layer
(``tssep_b200/torch_ops.py``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib, torch_ops
from ._lib import EPI_BF16, EPI_F32, EPI_HEAD


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def g_dtype() -> torch.dtype:
    """Storage type of the input projections G between the GEMM and the recurrence
    (TSSEP_G_DTYPE=bf16|f32, default bf16).  bf16 halves the largest intermediate of the path (2.9 GB per
    meeting and layer in f32) and is what the tensor-memory recurrence consumes (G is an MMA operand there);
    f32 routes every recurrence through the register-resident kernel (csrc/lstm.cu) -- a parity-test knob that
    separates the rounding of G from everything else."""
    return torch.float32 if os.environ.get("TSSEP_G_DTYPE", "bf16") == "f32" else torch.bfloat16


def operand_ld(cols: int) -> int:
    """Leading dimension (elements) of a bf16 GEMM operand with ``cols`` columns: rows of 128 bytes or more start
    on 128-byte boundaries, so that a 64-column TMA box row never straddles an extra L2 line (K = 513 rows of
    1040 bytes cost 13 % more L2 -> SM traffic than rows of 1152 bytes; measured on the birnn0 input projection)."""
    return round_up(cols, 64) if cols >= 64 else round_up(cols, 8)


def _gemm_impl() -> int:
    return 1 if os.environ.get("TSSEP_GEMM_IMPL", "tcgen05") == "simt" else 0


def gemm(A, lda, B, ldb, M, N, K, out, *, mode, ldo=0, batch=1, a_stride=0, a_div=1, b_stride=0, b_mod=None,
         bias=None, bias_stride=0, alpha=1.0, act=0, out_stride=0, out_div=None, out_stride_hi=0, mask=None,
         plane_map=None, n_blocks=0, row_len=0, impl=None, max_ctas=0):
    """``tssep_gemm``: out[z] = act(alpha * A[z / a_div] . B[z % b_mod]^T + bias[z % b_mod])."""
    torch_ops.op.gemm(A, lda, a_stride, a_div, B, ldb, b_stride, batch if b_mod is None else b_mod, bias, bias_stride,
                      M, N, K, batch, float(alpha), act, mode, out, ldo, out_stride, batch if out_div is None else out_div,
                      out_stride_hi, mask, plane_map, n_blocks, row_len, _gemm_impl() if impl is None else impl, max_ctas)


def row_pitch_view(t: torch.Tensor):
    """(tensor, pitch) of t (..., T, F) float32 for the kernels that take a row pitch: t itself when its rows are
    ``pitch >= F`` floats apart and everything in front of them is dense (the padded mask / logit buffers of the head
    GEMM, and every contiguous tensor), else a contiguous copy."""
    F = t.shape[-1]
    if t.dtype == torch.float32 and t.dim() >= 2 and (F == 1 or t.stride(-1) == 1):
        if t.is_contiguous():
            return t, F
        pitch = t.stride(-2)
        if pitch >= F:
            expect, ok = pitch * t.shape[-2], True
            for size, stride in zip(reversed(t.shape[:-2]), reversed(t.stride()[:-2])):
                if size != 1 and stride != expect:
                    ok = False
                    break
                expect *= size
            if ok:
                return t, pitch
    return t.float().contiguous(), F


def cast_bf16(src: torch.Tensor, ld_dst: int = None) -> torch.Tensor:
    """(rows, cols) f32 -> (rows, ld_dst) bf16 with zero padded columns."""
    _lib.require_cuda(src)
    src = src.contiguous()
    rows, cols = src.shape
    ld_dst = operand_ld(cols) if ld_dst is None else ld_dst
    dst = torch.empty((rows, ld_dst), dtype=torch.bfloat16, device=src.device)
    torch_ops.op.cast_bf16(src, rows, cols, cols, dst, ld_dst)
    return dst


def pack_whh(w_fwd: torch.Tensor, w_bwd: torch.Tensor, U: int, Up: int) -> torch.Tensor:
    _lib.require_cuda(w_fwd, w_bwd)
    n = 2 * (Up // 4) * (Up // 16) * 128
    out = torch.empty(n, dtype=torch.int32, device=w_fwd.device)
    torch_ops.op.pack_whh(w_fwd.contiguous(), w_bwd.contiguous(), U, Up, out)
    return out


def blstm_recurrence(G: torch.Tensor, wfrag: torch.Tensor, rows: int, T: int, Up: int, cluster: int = None,
                     fast_math: bool = None) -> torch.Tensor:
    """G (rows, T, 2, 4, Up) f32 or bf16 -> H (rows, T, 2*Up) bf16."""
    _lib.require_cuda(G, wfrag)
    if cluster is None:
        cluster = int(os.environ.get("TSSEP_LSTM_CLUSTER", "0"))
    if fast_math is None:
        fast_math = fast_math_default()
    H = torch.empty((rows, T, 2 * Up), dtype=torch.bfloat16, device=G.device)
    torch_ops.op.blstm_recurrence(G, int(G.dtype == torch.bfloat16), wfrag, H, rows, T, Up, cluster, int(fast_math))
    return H


def fast_math_default() -> bool:
    """Gate arithmetic of the recurrences: tanh.approx.f32 (max relative error 2^-11) unless TSSEP_LSTM_FAST_MATH=0
    selects the exp-based sigmoid / tanh."""
    return os.environ.get("TSSEP_LSTM_FAST_MATH", "1") == "1"


def pack_whh_ts(w_fwd: torch.Tensor, w_bwd: torch.Tensor, U: int, Up: int) -> torch.Tensor:
    """Tensor-memory image of W_hh for the TMEM-resident recurrence (bf16 pairs per 32-bit column)."""
    _lib.require_cuda(w_fwd, w_bwd)
    c = (Up + 63) // 64
    out = torch.empty(2 * c * 2 * (Up // 16) * 128 * 8, dtype=torch.int32, device=w_fwd.device)
    torch_ops.op.pack_whh_ts(w_fwd.contiguous(), w_bwd.contiguous(), U, Up, out)
    return out


def blstm_recurrence_ts(G: torch.Tensor, wimg: torch.Tensor, rows: int, T: int, Up: int,
                        fast_math: bool = None, rows_per_cluster: int = 0, k_split: int = -1,
                        tiles_per_cta: int = 0, sub_batches: int = 0) -> torch.Tensor:
    """Tensor-memory recurrence: G (rows, T, 2, 4, Up) bf16 -> H (rows, T, 2*Up) bf16."""
    _lib.require_cuda(G, wimg)
    if G.dtype != torch.bfloat16:
        raise TypeError(f"the tensor-memory recurrence consumes bf16 input projections, got {G.dtype}")
    if fast_math is None:
        fast_math = fast_math_default()
    H = torch.empty((rows, T, 2 * Up), dtype=torch.bfloat16, device=G.device)
    torch_ops.op.blstm_recurrence_ts(G, wimg, H, rows, T, Up, rows_per_cluster, tiles_per_cta, sub_batches, int(fast_math),
                                     k_split)
    return H


def recurrence_ts_capacity(Up: int, rows_per_cluster: int = 16, tiles_per_cta: int = 2, sub_batches: int = None) -> int:
    """Batch rows one launch of the tensor-memory recurrence advances in a single wave of clusters (``sub_batches``
    None: 2 for 64 rows per cluster, which exist only as two sub-batches, else 1; the capacity of a shape does not
    depend on it otherwise -- the clusters are the same size)."""
    if sub_batches is None:
        sub_batches = 2 if rows_per_cluster == 64 else 1
    n = _lib.load().tssep_blstm_recurrence_ts_capacity(Up, rows_per_cluster, tiles_per_cta, sub_batches)
    if n < 0:
        _lib.check(n, "tssep_blstm_recurrence_ts_capacity")
    return n


def instance_norm(x: torch.Tensor, unbiased=False, dim: int = -1, mode: int = 0) -> torch.Tensor:
    """Normalisation along ``dim`` (``tssep_instance_norm``): mode 0 (x - mean) / std, 1 x - mean, 2 x / rms."""
    _lib.require_cuda(x)
    x = x.contiguous().float()
    dim = dim % x.dim()
    cols = x.shape[dim]
    outer = 1
    for n in x.shape[:dim]:
        outer *= n
    inner = 1
    for n in x.shape[dim + 1:]:
        inner *= n
    out = torch.empty_like(x)
    torch_ops.op.instance_norm(x, outer, cols, inner, mode, int(bool(unbiased)), out)
    return out


__all__ = ["gemm", "cast_bf16", "operand_ld", "pack_whh", "blstm_recurrence", "pack_whh_ts", "blstm_recurrence_ts",
           "recurrence_ts_capacity", "instance_norm", "round_up", "fast_math_default", "EPI_F32", "EPI_BF16", "EPI_HEAD"]


def pcm16(x: torch.Tensor, scale: float = 32767.0, out: torch.Tensor = None) -> torch.Tensor:
    """float32 audio -> int16 PCM on the device (round to nearest, saturate); ``scale`` = 32767 / peak."""
    _lib.require_cuda(x)
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    torch_ops.op.pcm16(x, x.numel(), float(scale), out)
    return out
