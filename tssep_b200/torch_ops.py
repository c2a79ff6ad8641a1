"""The torch custom-op layer over the C ABI: ``torch.ops.tssep_b200.<entry point>``.

One operator per compute entry point of ``include/tssep_b200.h``, same argument order: device pointers become
``Tensor`` arguments (the operator passes ``data_ptr()`` on), outputs are caller-allocated tensors declared as mutated
(``Tensor(a!)``), scalars stay scalars, and the stream argument is the current stream of the first tensor's device.
Registered for the CUDA dispatch key ONLY -- a CPU tensor has no kernel to land on (there is no CPU fallback) -- plus a
fake implementation (no outputs to describe: every result is written into a caller-allocated tensor), so
``torch.compile`` / ``make_fx`` trace through the operators without graph breaks and without running them.

SURVEY.md §8b names this layer ("wrapped in torch.library custom ops so shapes propagate"); the ctypes binding
(``tssep_b200/_lib.py``) stays underneath as the only place that touches the shared library.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

NAMESPACE = "tssep_b200"

# name -> (schema of the operator, C argument plan).  Plan entries: a schema argument name (tensor -> pointer, scalar ->
# value) in C order; the trailing stream argument is implied.
_OPS = {
    "stft": (
        "(Tensor audio, int n_signals, int num_samples, Tensor window, Tensor twiddle, int size, int shift, "
        "int window_length, int fading, int T, Tensor(a!) X) -> ()",
        "tssep_stft"),
    "feature_stats": (
        "(Tensor X, int n_items, int x_item_stride, int T, int F, Tensor? mel_t, Tensor? mel_lo, Tensor? mel_hi, "
        "int n_mels, Tensor(a!) absmax_key, Tensor(b!) maxdb_key, Tensor(c!)? meldb) -> ()",
        "tssep_feature_stats"),
    "feature_write": (
        "(Tensor X, int n_items, int x_item_stride, int T, int F, Tensor absmax_key, Tensor maxdb_key, Tensor? meldb, "
        "Tensor? dct, int n_mels, int n_mfcc, int with_log1p, float top_db, int couple_batch, Tensor(a!)? feat_f32, "
        "Tensor(b!)? feat_bf16, int ld_bf16) -> ()",
        "tssep_feature_write"),
    "wpe": (
        "(Tensor Y, int D, int T, int F, int taps, int delay, int iterations, int psd_context, int statistics_mode, "
        "Tensor(a!) X, Tensor(b!) workspace, int workspace_bytes) -> ()",
        "tssep_wpe"),
    "pcm16": ("(Tensor x, int n, float scale, Tensor(a!) out) -> ()", "tssep_pcm16"),
    "log1p_abs": ("(Tensor X, int n, Tensor(a!) out) -> ()", "tssep_log1p_abs"),
    "ipd": (
        "(Tensor X, int lead, int D, int TF, Tensor second_channel, Tensor(a!) cos_out, Tensor(b!) sin_out) -> ()",
        "tssep_ipd"),
    "cast_bf16": (
        "(Tensor src, int rows, int cols, int ld_src, Tensor(a!) dst, int ld_dst) -> ()",
        "tssep_cast_bf16"),
    "instance_norm": (
        "(Tensor src, int outer, int cols, int inner, int mode, int unbiased, Tensor(a!) dst) -> ()",
        "tssep_instance_norm"),
    "fold_embedding": (
        "(int mode, Tensor W, int ldw, Tensor b, Tensor e, int Z, int N, int F, int A, Tensor(a!)? Wk, int ld_wk, "
        "Tensor(b!) bias_k) -> ()",
        "tssep_fold_embedding"),
    "head_expand_t": (
        "(Tensor small, int Z, int T, int n_blocks, int F, Tensor plane_map, Tensor(a!)? logit, Tensor(b!)? mask) -> ()",
        "tssep_head_expand_t"),
    "blstm_recurrence": (
        "(Tensor G, int g_dtype, Tensor Wfrag, Tensor(a!) H, int rows, int T, int Up, int cluster, int fast_math) -> ()",
        "tssep_blstm_recurrence"),
    "pack_whh": (
        "(Tensor whh_fwd, Tensor whh_bwd, int U, int Up, Tensor(a!) Wfrag) -> ()",
        "tssep_pack_whh"),
    "blstm_recurrence_ts": (
        "(Tensor G, Tensor Wimg, Tensor(a!) H, int rows, int T, int Up, int rows_per_cluster, int tiles_per_cta, int sub_batches, "
        "int gate_math, int k_split) -> ()",
        "tssep_blstm_recurrence_ts"),
    "pack_whh_ts": (
        "(Tensor whh_fwd, Tensor whh_bwd, int U, int Up, Tensor(a!) Wimg) -> ()",
        "tssep_pack_whh_ts"),
    "blstm_recurrence_train": (
        "(Tensor G, Tensor Wimg, Tensor(a!) H, Tensor(b!) gates, Tensor(c!) cstate, int rows, int T, int Up, "
        "int rows_per_cluster, int gate_math) -> ()",
        "tssep_blstm_recurrence_train"),
    "blstm_recurrence_bwd": (
        "(Tensor gates, Tensor cstate, Tensor dH, Tensor WTimg, Tensor(a!) dG, int rows, int T, int Up) -> ()",
        "tssep_blstm_recurrence_bwd"),
    "pack_whh_bwd": (
        "(Tensor whh_fwd, Tensor whh_bwd, int U, int Up, Tensor(a!) WTimg) -> ()",
        "tssep_pack_whh_bwd"),
    "mask_istft": (
        "(Tensor X, int x_item_stride, Tensor? mask, int mask_pitch, int Z, int n_spk, int T, int size, int shift, int window_length, "
        "int fading, Tensor synwin, Tensor twiddle, Tensor(a!)? stft_estimate, Tensor(b!)? time, int num_samples, "
        "Tensor(c!)? activity) -> ()",
        "tssep_mask_istft"),
    "bf_psd": (
        "(Tensor Y, Tensor mask, int Z, int K, int nmask, int D, int T, int F, Tensor(a!) psd) -> ()",
        "tssep_bf_psd"),
    "bf_mvdr_souden": (
        "(Tensor psd, int Z, int K, int nmask, int D, int F, int reference_channel, float eps, Tensor(a!) w) -> ()",
        "tssep_bf_mvdr_souden"),
    "bf_apply": (
        "(Tensor Y, Tensor w, Tensor? mask, int Z, int K, int nmask, int D, int T, int F, float masking_eps, "
        "Tensor(a!) out) -> ()",
        "tssep_bf_apply"),
    "activity": (
        "(Tensor mask, int n, int T, int F, int mask_pitch, Tensor(a!) activity) -> ()",
        "tssep_activity"),
    "median_threshold": (
        "(Tensor activity, int n, int T, int width, float threshold, Tensor(a!)? smooth, Tensor(b!)? active) -> ()",
        "tssep_median_threshold"),
    "segments": (
        "(Tensor active, int n, int T, int window_length, int shift, int fading, int num_samples, Tensor(a!) segments, "
        "Tensor(b!) counts, int max_segments, int index_mode) -> ()",
        "tssep_segments"),
    "stft_vad": (
        "(Tensor vad, int n, int num_samples, int window_length, int shift, int fading, int T, Tensor(a!) frames) -> ()",
        "tssep_stft_vad"),
    # tssep_gemm takes a descriptor struct: the operator carries its fields as arguments
    "gemm": (
        "(Tensor A, int lda, int a_stride, int a_div, Tensor B, int ldb, int b_stride, int b_mod, Tensor? bias, "
        "int bias_stride, int M, int N, int K, int batch, float alpha, int act, int mode, Tensor(a!)? out, int ldo, "
        "int out_stride, int out_div, int out_stride_hi, Tensor(b!)? mask, Tensor? plane_map, int n_blocks, int row_len, "
        "int impl, int max_ctas) -> ()",
        None),
}

_lib_def = torch.library.Library(NAMESPACE, "DEF")
_registered = False


def _first_tensor(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            return a
    raise RuntimeError("tssep_b200 operator called without a tensor argument")


# timeline labels of bench.py for the calls whose cost depends on their extent
_DETAIL = {
    "tssep_blstm_recurrence_ts": lambda a: f"rows={a[3]} T={a[4]}",
    "tssep_blstm_recurrence": lambda a: f"rows={a[4]} T={a[5]} regs",
    "tssep_blstm_recurrence_train": lambda a: f"rows={a[5]} T={a[6]} train",
    "tssep_blstm_recurrence_bwd": lambda a: f"rows={a[5]} T={a[6]} bwd",
}


def _make_impl(c_name):
    detail = _DETAIL.get(c_name)

    def impl(*args):
        tensors = [a for a in args if isinstance(a, torch.Tensor)]
        _lib.require_cuda(*tensors)
        c_args = [(_lib.ptr(a) if (a is None or isinstance(a, torch.Tensor)) else a) for a in args]
        _lib.call(c_name, *c_args, _lib.stream_of(_first_tensor(args)), detail=detail(args) if detail else None)

    return impl


def _gemm_impl(A, lda, a_stride, a_div, B, ldb, b_stride, b_mod, bias, bias_stride, M, N, K, batch, alpha, act, mode, out,
               ldo, out_stride, out_div, out_stride_hi, mask, plane_map, n_blocks, row_len, impl, max_ctas):
    _lib.require_cuda(A, B, out, bias, mask, plane_map)
    d = _lib.GemmDesc()
    d.A, d.lda, d.a_stride, d.a_div = A.data_ptr(), lda, a_stride, a_div
    d.B, d.ldb, d.b_stride, d.b_mod = B.data_ptr(), ldb, b_stride, b_mod
    d.bias, d.bias_stride = _lib.ptr(bias), bias_stride
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.alpha, d.act, d.mode = alpha, act, mode
    d.out, d.ldo, d.out_stride, d.out_div, d.out_stride_hi = _lib.ptr(out), ldo, out_stride, out_div, out_stride_hi
    d.mask, d.plane_map, d.n_blocks, d.row_len = _lib.ptr(mask), _lib.ptr(plane_map), n_blocks, row_len
    d.impl, d.max_ctas = impl, max_ctas
    _lib.call("tssep_gemm", C.byref(d), _lib.stream_of(A), detail=f"M={M} N={N} K={K} batch={batch} mode={mode}")


def _fake(*args, **kwargs):
    return None


def register():
    """Defines the operators (idempotent).  Imported for its side effect by ``tssep_b200/__init__.py``."""
    global _registered
    if _registered:
        return
    for name, (schema, c_name) in _OPS.items():
        _lib_def.define(name + schema)
        impl = _gemm_impl if c_name is None else _make_impl(c_name)
        _lib_def.impl(name, impl, "CUDA")
        torch.library.register_fake(f"{NAMESPACE}::{name}", _fake, lib=_lib_def)
    _registered = True


register()
op = getattr(torch.ops, NAMESPACE)

# the compute entry points of the header that have an operator (host-side queries have none)
WRAPPED_SYMBOLS = sorted(c if c else "tssep_gemm" for _, c in _OPS.values())
