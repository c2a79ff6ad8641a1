"""Drop-in ``RNNP_packed`` (tssep/train/rnnp.py:12-173): projected BLSTM stack.

Same constructor arguments, module list (so ``repr`` and ``state_dict`` keys
``net.0.weight_ih_l0`` ... match the reference) and ``forward`` contract for
2-/3-/4-D inputs.  The ``torch.nn.LSTM`` / ``Linear`` children only hold the
parameters; ``forward`` runs

    x . W_ih^T (+ b_ih + b_hh)      tcgen05 GEMM, both directions at once
    time recurrence                 persistent cluster kernel: W_hh resident in tensor memory (csrc/lstm_ts.cu)
                                    from 17 batch rows on, in registers (csrc/lstm.cu) below
    [h_fwd | h_bwd] . W_proj^T + b  tcgen05 GEMM with fused bias (+ tanh)

on bf16 operands with fp32 accumulation and fp32 cell state.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops

# Recurrence kernels (all compute the same operator):
#   regs  csrc/lstm.cu     W_hh in registers, mma.sync, rows in the plain (row, t) layout
#   ts    csrc/lstm_ts.cu  W_hh in tensor memory, tcgen05.mma with A from TMEM, rows ordered (group, t, b32)
#   tc    csrc/lstm_tc.cu  W_hh in shared memory, tcgen05.mma (kept for A/B measurements)
# TSSEP_LSTM_KERNEL=regs|ts|tc|auto; auto takes the tensor-memory kernel from TS_MIN_ROWS batch rows on.
TS_MIN_ROWS = int(os.environ.get("TSSEP_TS_MIN_ROWS", "17"))


def rec_kernel(rows: int) -> str:
    choice = os.environ.get("TSSEP_LSTM_KERNEL", "auto")
    if choice == "auto":
        return "ts" if rows >= TS_MIN_ROWS else "regs"
    return choice


def use_tc_recurrence(rows: int) -> bool:
    """True when the recurrence runs on a tcgen05 kernel, i.e. on rows ordered (group, t, b32)."""
    return rec_kernel(rows) in ("ts", "tc")


class LayerPack:
    """Device-resident, kernel-ready copies of one BLSTM + projection layer (a derived cache)."""

    def __init__(self, lstm: torch.nn.LSTM, linear: torch.nn.Linear):
        U, I = lstm.hidden_size, lstm.input_size
        Up = ops.round_up(U, 16)
        if Up > 320:
            raise NotImplementedError(f"hidden size {U} > 320 is not supported by the recurrence kernel yet")
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
            self.whh = ops.pack_whh(self._whh_f32[0], self._whh_f32[1], U, Up)
            self.whh_tc = self.whh_ts = None  # built on first use by the tcgen05 recurrences
            self.w_ih_tc = self.bias_tc = None
            wp = torch.zeros((self.hdim, 2 * Up), dtype=torch.float32, device=dev)
            wp[:, :U] = linear.weight.detach().float()[:, :U]
            wp[:, Up:Up + U] = linear.weight.detach().float()[:, U:]
            self.w_proj = ops.cast_bf16(wp, 2 * Up)
            self.b_proj = linear.bias.detach().float().contiguous()

    # -- the three stages ------------------------------------------------------
    def input_gemm(self, xb: torch.Tensor, ld: int, rows_t: int) -> torch.Tensor:
        """xb (rows*T, ld) bf16 -> G (rows*T, 8*Up) f32."""
        gd = ops.g_dtype()
        G = torch.empty((rows_t, 8 * self.Up), dtype=gd, device=xb.device)
        ops.gemm(xb, ld, self.w_ih, self.ld_in, rows_t, 8 * self.Up, self.I, G,
                 mode=ops.EPI_BF16 if gd == torch.bfloat16 else ops.EPI_F32, ldo=8 * self.Up, bias=self.bias)
        return G

    def recurrence(self, G: torch.Tensor, rows: int, T: int) -> torch.Tensor:
        """G (rows, T, 8Up) -> H (rows, T, 2Up): the tensor-memory kernel (csrc/lstm_ts.cu, row layout) from
        TS_MIN_ROWS rows on when G is bf16, else the register-resident mma.sync kernel (csrc/lstm.cu)."""
        if rec_kernel(rows) == "ts" and G.dtype == torch.bfloat16:
            if self.whh_ts is None:
                self.whh_ts = ops.pack_whh_ts(self._whh_f32[0], self._whh_f32[1], self.U, self.Up)
            return ops.blstm_recurrence_ts(G, self.whh_ts, rows, T, self.Up, layout="rows")
        return ops.blstm_recurrence(G, self.whh, rows, T, self.Up)

    # -- throughput path: rows ordered (group, t, b32), weights in shared memory, tcgen05 -------------
    def input_gemm_bt(self, xb: torch.Tensor, ld: int, mrows: int, kdim: int = None) -> torch.Tensor:
        """xb (groups*T*32, ld) bf16 -> G (groups, T, 8Up, 32) f32 (batch row innermost)."""
        if self.w_ih_tc is None:
            # column order of the tcgen05 recurrence: n = dir*4Up + (unit/8)*32 + (unit%8)*4 + gate
            Up = self.Up
            w = self.w_ih_f32.view(2, 4, Up // 8, 8, self.I).permute(0, 2, 3, 1, 4).reshape(8 * Up, self.I)
            self.w_ih_tc = ops.cast_bf16(w.contiguous(), self.ld_in)
            self.bias_tc = self.bias.view(2, 4, Up // 8, 8).permute(0, 2, 3, 1).reshape(8 * Up).contiguous()
        gd = ops.g_dtype()
        G = torch.empty((mrows * 8 * self.Up,), dtype=gd, device=xb.device)
        ops.gemm(xb, ld, self.w_ih_tc, self.ld_in, mrows, 8 * self.Up, self.I if kdim is None else kdim, G,
                 mode=ops.EPI_BF16_BT if gd == torch.bfloat16 else ops.EPI_F32_BT, bias=self.bias_tc)
        return G

    def recurrence_tc(self, G: torch.Tensor, rows: int, T: int) -> torch.Tensor:
        """tcgen05 kernels (csrc/lstm_ts.cu, csrc/lstm_tc.cu): H (groups*T*32, 2Up), rows (group, t, b)."""
        if rec_kernel(rows) == "tc":
            if self.whh_tc is None:
                self.whh_tc = ops.pack_whh_tc(self._whh_f32[0], self._whh_f32[1], self.U, self.Up)
            return ops.blstm_recurrence_tc(G, self.whh_tc, rows, T, self.Up)
        if self.whh_ts is None:
            self.whh_ts = ops.pack_whh_ts(self._whh_f32[0], self._whh_f32[1], self.U, self.Up)
        return ops.blstm_recurrence_ts(G, self.whh_ts, rows, T, self.Up)

    def projection(self, H: torch.Tensor, rows_t: int, out: torch.Tensor, *, mode: int, ldo: int, act: int,
                   batch=1, a_stride=0, M=None, out_stride=0, out_div=None, out_stride_hi=0, row_map=None):
        """H (rows*T, 2*Up) bf16 -> out (bias and optional tanh fused)."""
        ops.gemm(H, 2 * self.Up, self.w_proj, 2 * self.Up, rows_t if M is None else M, self.hdim, 2 * self.Up, out,
                 mode=mode, ldo=ldo, bias=self.b_proj, act=act, batch=batch, a_stride=a_stride, b_mod=1,
                 out_stride=out_stride, out_div=out_div, out_stride_hi=out_stride_hi, row_map=row_map)


def param_key(module: torch.nn.Module):
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
        x = xs_pack.reshape(rows * T, D).float()
        packs = self.layer_packs()
        # TSSEP_TS_LAYOUT=bt routes this generic entry point through the (group, t, b32) row order the mask
        # estimator uses for its speaker-independent layers; by default the tensor-memory kernel reads and
        # writes plain (row, t) layouts here and no re-ordering copies are needed.
        if rec_kernel(rows) == "tc" or (rec_kernel(rows) == "ts" and os.environ.get("TSSEP_TS_LAYOUT", "rows") == "bt"):
            return self._forward_tc(x, rows, T, D, packs).reshape(*shape[:-1], packs[-1].hdim)
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

    def _forward_tc(self, x, rows, T, D, packs):
        """Same stack through the tcgen05 recurrence.  The kernels want rows ordered (group, t, b32);
        this generic entry point re-orders with torch copies (MaskEstimator_v2 produces that order
        directly and never takes this route)."""
        groups = (rows + 31) // 32
        mrows = groups * T * 32
        xp = torch.zeros((groups * 32, T, D), dtype=torch.float32, device=x.device)
        xp[:rows] = x.view(rows, T, D)
        xb = ops.cast_bf16(xp.view(groups, 32, T, D).permute(0, 2, 1, 3).reshape(mrows, D))
        ld = ops.operand_ld(D)
        out = None
        for li, pk in enumerate(packs):
            G = pk.input_gemm_bt(xb, ld, mrows)
            H = pk.recurrence_tc(G, rows, T)
            del G
            if li == len(packs) - 1:
                out = torch.empty((mrows, pk.hdim), dtype=torch.float32, device=x.device)
                pk.projection(H, mrows, out, mode=ops.EPI_F32, ldo=pk.hdim, act=0)
            else:
                ld = ops.operand_ld(pk.hdim)
                xb = torch.empty((mrows, ld), dtype=torch.bfloat16, device=x.device)
                pk.projection(H, mrows, xb, mode=ops.EPI_BF16, ldo=ld, act=1)
            del H
        out = out.view(groups, T, 32, -1).permute(0, 2, 1, 3).reshape(groups * 32, T, -1)[:rows]
        return out.reshape(rows * T, -1)
