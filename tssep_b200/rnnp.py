"""Drop-in ``RNNP_packed`` (tssep/train/rnnp.py:12-173): projected BLSTM stack.

Same constructor arguments, module list (so ``repr`` and ``state_dict`` keys
``net.0.weight_ih_l0`` ... match the reference) and ``forward`` contract for
2-/3-/4-D inputs.  The ``torch.nn.LSTM`` / ``Linear`` children only hold the
parameters; ``forward`` runs

    x . W_ih^T (+ b_ih + b_hh)      tcgen05 GEMM, both directions at once
    time recurrence                 persistent cluster kernel, W_hh resident in tensor memory (csrc/lstm_ts.cu)
    [h_fwd | h_bwd] . W_proj^T + b  tcgen05 GEMM with fused bias (+ tanh)

on bf16 operands with fp32 accumulation and fp32 cell state.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops, torch_ops

# Recurrence kernels (both compute the same operator on the plain (row, t) layouts):
#   ts    csrc/lstm_ts.cu  W_hh in tensor memory, tcgen05.mma, bf16 G: the product path for every row count
#   regs  csrc/lstm.cu     W_hh in registers, mma.sync, f32 or bf16 G (Up <= 320): parity-test variant
# TSSEP_LSTM_KERNEL=auto|ts|regs; auto = ts whenever G is bf16 (the default storage type, ops.g_dtype).
TS_MAX_UP = 384


def rec_kernel(g_dtype: torch.dtype = torch.bfloat16, Up: int = 0) -> str:
    choice = os.environ.get("TSSEP_LSTM_KERNEL", "auto")
    if choice == "auto":
        return "ts" if g_dtype == torch.bfloat16 and Up <= TS_MAX_UP else "regs"
    return choice


# Cluster shapes of the tensor-memory recurrence as (row tiles per CTA, rows per cluster, sub-batches, us per dependent
# step at U = 300 on B200): the best shape per cluster width of csrc/lstm_ts.cu::kTsShapes.
TS_SHAPES = ((1, 8, 1, 0.78), (2, 8, 1, 0.85), (2, 16, 2, 1.0), (2, 32, 2, 1.54), (2, 64, 2, 3.06))

_cta_budget = None


def set_cta_budget(n):
    """Upper bound on the CTAs (= SMs: one CTA per SM) a recurrence launch may occupy; ``None`` = the library's choice
    (the fastest shape that fits the device).  A serving loop that keeps two steps in flight on two streams gives
    every step half of the SMs, so that the latency-bound recurrences of both run side by side instead of one waiting
    for the clusters of the other to drain (bench.py at <= 16 meetings per GPU)."""
    global _cta_budget
    _cta_budget = None if n is None else int(n)


def choose_ts_shape(rows: int, Up: int, cta_budget: int, capacity=None):
    """(rows_per_cluster, tiles_per_cta, sub_batches) of the cheapest shape whose launch needs at most ``cta_budget``
    CTAs, or (0, 0, 0) -- let the library choose -- when none does.  ``capacity(Up, rows_per_cluster, tiles, subs)``:
    rows one wave of co-resident clusters holds (default: the device query)."""
    capacity = ops.recurrence_ts_capacity if capacity is None else capacity
    best, pick = None, (0, 0, 0)
    for tiles, rpc, subs, cost in TS_SHAPES:
        ctas_per_cluster = 2 * ((Up + 63) // 64) // tiles
        if ctas_per_cluster > 16:
            continue
        ctas = 2 * (-(-rows // rpc)) * ctas_per_cluster           # both directions
        cap = capacity(Up, rpc, tiles, subs)
        if cap <= 0 or ctas > cta_budget:
            continue
        total = cost * (-(-rows // cap))
        if best is None or total < best:
            best, pick = total, (rpc, tiles, subs)
    return pick


class LayerPack:
    """Device-resident, kernel-ready copies of one BLSTM + projection layer (a derived cache)."""

    def __init__(self, lstm: torch.nn.LSTM, linear: torch.nn.Linear):
        U, I = lstm.hidden_size, lstm.input_size
        Up = ops.round_up(U, 16)
        if Up > TS_MAX_UP:
            raise NotImplementedError(f"hidden size {U} > {TS_MAX_UP} exceeds the tensor-memory budget of the recurrence "
                                      "kernel (include/tssep_b200.h, tssep_blstm_recurrence_ts)")
        dev = lstm.weight_ih_l0.device
        _lib.require_cuda(lstm.weight_ih_l0)
        self.U, self.Up, self.I = U, Up, I
        self.hdim = linear.out_features
        with torch.no_grad():
            w = torch.zeros((2, 4, Up, I), dtype=torch.float32, device=dev)
            w[0, :, :U] = lstm.weight_ih_l0.detach().float().view(4, U, I)
            w[1, :, :U] = lstm.weight_ih_l0_reverse.detach().float().view(4, U, I)
            self.w_ih_f32 = w.view(8 * Up, I)  # rows [dir][gate][unit]
            b = torch.zeros((2, 4, Up), dtype=torch.float32, device=dev)
            b[0, :, :U] = (lstm.bias_ih_l0 + lstm.bias_hh_l0).detach().float().view(4, U)
            b[1, :, :U] = (lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse).detach().float().view(4, U)
            self.bias = b.view(8 * Up).contiguous()
            self.ld_in = ops.operand_ld(I)
            self.w_ih = ops.cast_bf16(self.w_ih_f32, self.ld_in)
            self._whh_f32 = (lstm.weight_hh_l0.detach().float().contiguous(),
                             lstm.weight_hh_l0_reverse.detach().float().contiguous())
            self.whh_regs = self.whh_ts = None  # packed on first use by the respective kernel
            wp = torch.zeros((self.hdim, 2 * Up), dtype=torch.float32, device=dev)
            wp[:, :U] = linear.weight.detach().float()[:, :U]
            wp[:, Up:Up + U] = linear.weight.detach().float()[:, U:]
            self._w_proj_f32 = wp
            self.w_proj = ops.cast_bf16(wp, 2 * Up)
            self._w_proj_t = self._w_ih_t = self._whh_bwd = None  # backward operands, built on first use
            self.b_proj = linear.bias.detach().float().contiguous()

    # -- the three stages ------------------------------------------------------
    def input_gemm(self, xb: torch.Tensor, ld: int, rows_t: int) -> torch.Tensor:
        """xb (rows*T, ld) bf16 -> G (rows*T, 8*Up) bf16 (f32 with TSSEP_G_DTYPE=f32)."""
        gd = ops.g_dtype()
        G = torch.empty((rows_t, 8 * self.Up), dtype=gd, device=xb.device)
        ops.gemm(xb, ld, self.w_ih, self.ld_in, rows_t, 8 * self.Up, self.I, G,
                 mode=ops.EPI_BF16 if gd == torch.bfloat16 else ops.EPI_F32, ldo=8 * self.Up, bias=self.bias)
        return G

    def whh_ts_image(self) -> torch.Tensor:
        if self.whh_ts is None:
            self.whh_ts = ops.pack_whh_ts(self._whh_f32[0], self._whh_f32[1], self.U, self.Up)
        return self.whh_ts

    # -- operands of the backward pass (tssep_b200/autograd.py) ------------------------------------------------
    def whh_bwd_image(self) -> torch.Tensor:
        """Transposed tensor-memory image of W_hh for the BPTT kernel (csrc/lstm_bwd.cu)."""
        if self._whh_bwd is None:
            c = (self.Up + 63) // 64
            out = torch.empty(2 * c * ((c + 1) // 2) * 16 * 128 * 8, dtype=torch.int32, device=self.bias.device)
            torch_ops.op.pack_whh_bwd(self._whh_f32[0], self._whh_f32[1], self.U, self.Up, out)
            self._whh_bwd = out
        return self._whh_bwd

    def w_proj_t(self):
        """(2Up, P) bf16: B operand of dH = dproj . W_proj."""
        if self._w_proj_t is None:
            ld = ops.operand_ld(self.hdim)
            self._w_proj_t = (ops.cast_bf16(self._w_proj_f32.t().contiguous(), ld), ld)
        return self._w_proj_t

    def w_ih_t(self):
        """(I, 8Up) bf16 with columns ordered [dir][unit][gate] (the order of dG): B operand of dx = dG . W_ih."""
        if self._w_ih_t is None:
            Up = self.Up
            w = self.w_ih_f32.view(2, 4, Up, self.I).permute(0, 2, 1, 3).reshape(8 * Up, self.I)
            ld = ops.operand_ld(8 * Up)
            self._w_ih_t = (ops.cast_bf16(w.t().contiguous(), ld), ld)
        return self._w_ih_t

    def recurrence(self, G: torch.Tensor, rows: int, T: int) -> torch.Tensor:
        """G (rows, T, 8Up) -> H (rows, T, 2Up)."""
        if rec_kernel(G.dtype, self.Up) == "ts":
            self.whh_ts_image()
            # host-side tuning knobs, handed to the library as explicit arguments (0 / -1 = let it choose)
            rpc, tiles = int(os.environ.get("TSSEP_TS_ROWS", "0")), int(os.environ.get("TSSEP_TS_TILES", "0"))
            subs = int(os.environ.get("TSSEP_TS_SUBS", "0"))
            if _cta_budget is not None and (rpc, tiles, subs) == (0, 0, 0):
                rpc, tiles, subs = choose_ts_shape(rows, self.Up, _cta_budget)
            return ops.blstm_recurrence_ts(G, self.whh_ts, rows, T, self.Up, rows_per_cluster=rpc,
                                           k_split=int(os.environ.get("TSSEP_TS_KSPLIT", "-1")),
                                           tiles_per_cta=tiles, sub_batches=subs)
        if self.whh_regs is None:
            self.whh_regs = ops.pack_whh(self._whh_f32[0], self._whh_f32[1], self.U, self.Up)
        return ops.blstm_recurrence(G, self.whh_regs, rows, T, self.Up)

    def projection(self, H: torch.Tensor, rows_t: int, out: torch.Tensor, *, mode: int, ldo: int, act: int,
                   batch=1, a_stride=0, M=None, out_stride=0, out_div=None, out_stride_hi=0):
        """H (rows*T, 2*Up) bf16 -> out (bias and optional tanh fused)."""
        ops.gemm(H, 2 * self.Up, self.w_proj, 2 * self.Up, rows_t if M is None else M, self.hdim, 2 * self.Up, out,
                 mode=mode, ldo=ldo, bias=self.b_proj, act=act, batch=batch, a_stride=a_stride, b_mod=1,
                 out_stride=out_stride, out_div=out_div, out_stride_hi=out_stride_hi)


def param_key(module: torch.nn.Module):
    """Identity + version + storage of every parameter.  In-place edits THROUGH ``.data`` (``p.data.copy_()``) bump
    neither: call ``invalidate_caches()`` on the owning module after such an edit."""
    return tuple((id(p), p._version, p.data_ptr()) for p in module.parameters())


class RNNP_packed(torch.nn.Module):
    """RNN with projection layers; see the module docstring.

    >>> RNNP_packed(512, 2, 300, 320, 0)  # doctest: +NORMALIZE_WHITESPACE
    RNNP_packed(
      (net): ModuleList(
        (0): LSTM(512, 300, batch_first=True, bidirectional=True)
        (1): Linear(in_features=600, out_features=320, bias=True)
        (2): Dropout(p=0, inplace=False)
        (3): Tanh()
        (4): LSTM(320, 300, batch_first=True, bidirectional=True)
        (5): Linear(in_features=600, out_features=320, bias=True)
      )
    )
    """

    def __init__(self, idim, elayers, cdim, hdim, dropout, typ="blstm", return_states=False):
        super().__init__()
        if typ != "blstm":
            raise NotImplementedError(f"typ={typ!r}: only 'blstm' (the value used by MaskEstimator_v2, "
                                      "tssep/train/net.py:545-552) is implemented")
        net = []
        for i in range(elayers):
            net.append(torch.nn.LSTM(idim if i == 0 else hdim, cdim, num_layers=1, bidirectional=True,
                                     batch_first=True))
            net.append(torch.nn.Linear(2 * cdim, hdim))
            if i < elayers - 1:
                net.append(torch.nn.Dropout(p=dropout))
                net.append(torch.nn.Tanh())
        self.net = torch.nn.ModuleList(net)
        self.elayers, self.cdim, self.typ, self.bidir = elayers, cdim, typ, True
        self.idim, self.hdim = idim, hdim
        self.dropout, self.return_states = dropout, return_states
        self._packs, self._packs_key = None, None

    def invalidate_caches(self):
        """Drops the packed weight copies (needed only after in-place edits through ``.data``, see ``param_key``)."""
        self._packs, self._packs_key = None, None

    def layer_packs(self):
        """Kernel-ready weights; rebuilt whenever a parameter was modified, moved or reloaded."""
        key = param_key(self)
        if self._packs is None or self._packs_key != key:
            mods = list(self.net)
            pairs = [(m, mods[i + 1]) for i, m in enumerate(mods) if isinstance(m, torch.nn.LSTM)]
            self._packs = [LayerPack(lstm, lin) for lstm, lin in pairs]
            self._packs_key = key
        return self._packs

    def forward(self, xs_pack, prev_state=None):
        assert prev_state is None, prev_state
        if self.return_states:
            raise NotImplementedError("return_states=True is not implemented (unused by the reference configs)")
        if self.training and self.dropout:
            raise NotImplementedError("dropout > 0 in training mode is not implemented on the CUDA path")
        if isinstance(xs_pack, torch.nn.utils.rnn.PackedSequence):
            raise NotImplementedError("PackedSequence input (unsupported by the reference too, rnnp.py:29-31)")
        _lib.require_cuda(xs_pack)
        shape = xs_pack.shape
        if len(shape) not in (2, 3, 4):
            raise KeyError(len(shape))
        T, D = shape[-2:]
        rows = 1
        for s in shape[:-2]:
            rows *= s
        if self.training and torch.is_grad_enabled():
            return self.forward_train(xs_pack.reshape(rows, T, D)).reshape(*shape[:-1], self.hdim)
        x = xs_pack.reshape(rows * T, D).float()
        packs = self.layer_packs()
        xb, ld = ops.cast_bf16(x), ops.operand_ld(D)
        out = None
        for li, pk in enumerate(packs):
            G = pk.input_gemm(xb, ld, rows * T)
            H = pk.recurrence(G, rows, T)
            del G
            if li == len(packs) - 1:
                out = torch.empty((rows * T, pk.hdim), dtype=torch.float32, device=x.device)
                pk.projection(H, rows * T, out, mode=ops.EPI_F32, ldo=pk.hdim, act=0)
            else:
                ld = ops.operand_ld(pk.hdim)
                xb = torch.empty((rows * T, ld), dtype=torch.bfloat16, device=x.device)
                pk.projection(H, rows * T, xb, mode=ops.EPI_BF16, ldo=ld, act=1)
            del H
        return out.reshape(*shape[:-1], packs[-1].hdim)

    def forward_train(self, x: torch.Tensor) -> torch.Tensor:
        """Differentiable forward (training mode): x (rows, T, idim) f32 -> (rows, T, hdim) f32 through
        ``tssep_b200.autograd.RNNPLayerFn`` (own kernels forward and backward, BPTT in csrc/lstm_bwd.cu)."""
        from .autograd import rnnp_layer

        packs = self.layer_packs()
        mods = list(self.net)
        pairs = [(m, mods[i + 1]) for i, m in enumerate(mods) if isinstance(m, torch.nn.LSTM)]
        h = x.float()
        for li, ((lstm, lin), pk) in enumerate(zip(pairs, packs)):
            h = rnnp_layer(lstm, lin, pk, h, act_tanh=li < len(packs) - 1)
        return h
