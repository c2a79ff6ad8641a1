"""Thin typed wrappers around the C-ABI calls shared by several modules.

Nothing here computes on the host: each function validates, allocates the
output with torch (device memory only) and enqueues one kernel from
``libtssep_b200.so`` on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import EPI_BF16, EPI_F32, EPI_HEAD, GemmDesc


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
    _lib.require_cuda(A, B, out, bias, mask, plane_map)
    d = GemmDesc()
    d.A, d.lda, d.a_stride, d.a_div = A.data_ptr(), lda, a_stride, a_div
    d.B, d.ldb, d.b_stride, d.b_mod = B.data_ptr(), ldb, b_stride, batch if b_mod is None else b_mod
    d.bias, d.bias_stride = _lib.ptr(bias), bias_stride
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.alpha, d.act, d.mode = alpha, act, mode
    d.out, d.ldo, d.out_stride, d.out_stride_hi = _lib.ptr(out), ldo, out_stride, out_stride_hi
    d.out_div = batch if out_div is None else out_div
    d.mask, d.plane_map, d.n_blocks, d.row_len = _lib.ptr(mask), _lib.ptr(plane_map), n_blocks, row_len
    d.impl = _gemm_impl() if impl is None else impl
    d.max_ctas = max_ctas
    _lib.call("tssep_gemm", C.byref(d), _lib.stream_of(A), detail=f"M={M} N={N} K={K} batch={batch} mode={mode}")


def cast_bf16(src: torch.Tensor, ld_dst: int = None) -> torch.Tensor:
    """(rows, cols) f32 -> (rows, ld_dst) bf16 with zero padded columns."""
    _lib.require_cuda(src)
    src = src.contiguous()
    rows, cols = src.shape
    ld_dst = operand_ld(cols) if ld_dst is None else ld_dst
    dst = torch.empty((rows, ld_dst), dtype=torch.bfloat16, device=src.device)
    _lib.call("tssep_cast_bf16", src.data_ptr(), rows, cols, cols, dst.data_ptr(), ld_dst, _lib.stream_of(src))
    return dst


def pack_whh(w_fwd: torch.Tensor, w_bwd: torch.Tensor, U: int, Up: int) -> torch.Tensor:
    _lib.require_cuda(w_fwd, w_bwd)
    n = 2 * (Up // 4) * (Up // 16) * 128
    out = torch.empty(n, dtype=torch.int32, device=w_fwd.device)
    _lib.call("tssep_pack_whh", w_fwd.contiguous().data_ptr(), w_bwd.contiguous().data_ptr(), U, Up, out.data_ptr(),
              _lib.stream_of(w_fwd))
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
    _lib.call("tssep_blstm_recurrence", G.data_ptr(), int(G.dtype == torch.bfloat16), wfrag.data_ptr(), H.data_ptr(),
              rows, T, Up, cluster,
              int(fast_math), _lib.stream_of(G))
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
    _lib.call("tssep_pack_whh_ts", w_fwd.contiguous().data_ptr(), w_bwd.contiguous().data_ptr(), U, Up,
              out.data_ptr(), _lib.stream_of(w_fwd))
    return out


def blstm_recurrence_ts(G: torch.Tensor, wimg: torch.Tensor, rows: int, T: int, Up: int,
                        fast_math: bool = None, rows_per_cluster: int = 0, k_split: int = -1) -> torch.Tensor:
    """Tensor-memory recurrence: G (rows, T, 2, 4, Up) bf16 -> H (rows, T, 2*Up) bf16."""
    _lib.require_cuda(G, wimg)
    if G.dtype != torch.bfloat16:
        raise TypeError(f"the tensor-memory recurrence consumes bf16 input projections, got {G.dtype}")
    if fast_math is None:
        fast_math = fast_math_default()
    H = torch.empty((rows, T, 2 * Up), dtype=torch.bfloat16, device=G.device)
    _lib.call("tssep_blstm_recurrence_ts", G.data_ptr(), wimg.data_ptr(), H.data_ptr(), rows, T, Up, rows_per_cluster,
              int(fast_math), k_split, _lib.stream_of(G), detail=f"rows={rows} T={T}")
    return H


def recurrence_ts_capacity(Up: int, rows_per_cluster: int = 16) -> int:
    """Batch rows one launch of the tensor-memory recurrence advances in a single wave of clusters."""
    n = _lib.load().tssep_blstm_recurrence_ts_capacity(Up, rows_per_cluster)
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
    _lib.call("tssep_instance_norm", x.data_ptr(), outer, cols, inner, mode, int(bool(unbiased)), out.data_ptr(),
              _lib.stream_of(x))
    return out


__all__ = ["gemm", "cast_bf16", "operand_ld", "pack_whh", "blstm_recurrence", "pack_whh_ts", "blstm_recurrence_ts",
           "recurrence_ts_capacity", "instance_norm", "round_up", "fast_math_default", "EPI_F32", "EPI_BF16", "EPI_HEAD"]
