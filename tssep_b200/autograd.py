"""Training step support (BASELINE config 5): differentiable RNNP layers and iSTFT on the B200 kernels.

The reference trains through ``torch.nn.LSTM`` / ``Linear`` autograd (tssep/train/rnnp.py:143-159) under
``LogMAE`` on the iSTFT output (tssep/train/loss.py:219-247, model.py:661-669).  Here one ``torch.autograd.Function``
covers a projected BLSTM layer:

    forward   x . W_ih^T            tssep_gemm (tcgen05)
              time recurrence       tssep_blstm_recurrence_train (tensor-memory kernel, stores gates and c_t)
              [h_f | h_b] . W_p^T   tssep_gemm (+ bias, + tanh)
    backward  d proj -> dH          tssep_gemm with the transposed projection weights
              BPTT                  tssep_blstm_recurrence_bwd (csrc/lstm_bwd.cu)
              dG -> dx              tssep_gemm with the transposed input weights
              weight gradients      plain GEMMs over (rows*T): cuBLAS through torch.mm (what the task statement
                                    reserves library GEMMs for), bias gradients: column sums

bf16 operands, f32 accumulation, f32 cell state and f32 parameter gradients (the "bf16" of config 5).
"""
from __future__ import annotations

import torch

from . import ops, torch_ops


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a (M, N)^T-free: returns a @ b in f32 for bf16 operands (cuBLAS, f32 accumulation and output)."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        # older torch: accumulate chunks of the long K dimension in f32
        out = torch.zeros((a.shape[0], b.shape[1]), dtype=torch.float32, device=a.device)
        step = 1 << 16
        for k0 in range(0, a.shape[1], step):
            out += torch.mm(a[:, k0:k0 + step], b[k0:k0 + step]).float()
        return out


class RNNPLayerFn(torch.autograd.Function):
    """y = act([h_fwd | h_bwd] W_p^T + b_p),  h = BLSTM(x) for x (rows, T, I) -> y (rows, T, P), f32 in / out."""

    @staticmethod
    def forward(ctx, x, w_ih, w_ih_r, w_hh, w_hh_r, b_ih, b_ih_r, b_hh, b_hh_r, w_proj, b_proj, pack, act_tanh):
        rows, T, I = x.shape
        U, Up, P = pack.U, pack.Up, pack.hdim
        dev = x.device
        xb = ops.cast_bf16(x.reshape(rows * T, I).float())
        ld = ops.operand_ld(I)
        G = torch.empty((rows * T, 8 * Up), dtype=torch.bfloat16, device=dev)
        ops.gemm(xb, ld, pack.w_ih, pack.ld_in, rows * T, 8 * Up, I, G, mode=ops.EPI_BF16, ldo=8 * Up, bias=pack.bias)
        H = torch.empty((rows, T, 2 * Up), dtype=torch.bfloat16, device=dev)
        gates = torch.empty((rows, T, 2, Up, 4), dtype=torch.bfloat16, device=dev)
        cstate = torch.empty((rows, T, 2, Up), dtype=torch.float32, device=dev)
        torch_ops.op.blstm_recurrence_train(G, pack.whh_ts_image(), H, gates, cstate, rows, T, Up, 0,
                                            int(ops.fast_math_default()))
        del G
        y = torch.empty((rows * T, P), dtype=torch.float32, device=dev)
        pack.projection(H.view(rows * T, 2 * Up), rows * T, y, mode=ops.EPI_F32, ldo=P, act=1 if act_tanh else 0)
        ctx.pack, ctx.act_tanh, ctx.dims = pack, act_tanh, (rows, T, I, U, Up, P, ld)
        ctx.save_for_backward(xb, H, gates, cstate, y if act_tanh else None)
        return y.view(rows, T, P)

    @staticmethod
    def backward(ctx, dy):
        xb, H, gates, cstate, y = ctx.saved_tensors
        pack = ctx.pack
        rows, T, I, U, Up, P, ld = ctx.dims
        M = rows * T
        dev = dy.device
        dproj = dy.reshape(M, P).float()
        if ctx.act_tanh:
            dproj = dproj * (1.0 - y * y)
        # projection: bias / weight gradients (library GEMM), dH through the tensor-core GEMM
        g_b_proj = dproj.sum(0)
        ldp = ops.operand_ld(P)
        dpb = ops.cast_bf16(dproj, ldp)
        H2 = H.view(M, 2 * Up)
        gw = _mm_f32(dpb[:, :P].t(), H2)  # (P, 2Up)
        g_w_proj = torch.cat([gw[:, :U], gw[:, Up:Up + U]], dim=1)
        dH = torch.empty((M, 2 * Up), dtype=torch.bfloat16, device=dev)
        wpt, ld_wpt = pack.w_proj_t()
        ops.gemm(dpb, ldp, wpt, ld_wpt, M, 2 * Up, P, dH, mode=ops.EPI_BF16, ldo=2 * Up, b_mod=1)
        del dpb, dproj
        # BPTT
        dG = torch.empty((rows, T, 2, Up, 4), dtype=torch.bfloat16, device=dev)
        torch_ops.op.blstm_recurrence_bwd(gates, cstate, dH, pack.whh_bwd_image(), dG, rows, T, Up)
        del dH
        dG2 = dG.view(M, 8 * Up)  # columns ordered [dir][unit][gate]
        # biases: both LSTM biases enter the pre-activation additively
        gb = dG2.float().sum(0).view(2, Up, 4).permute(0, 2, 1)[:, :, :U].reshape(2, 4 * U)
        # input weights: dW_ih[(dir, gate, unit), :] = sum_m dG[m, (dir, unit, gate)] x[m, :]
        gwi = _mm_f32(dG2.t(), xb[:, :I]).view(2, Up, 4, I).permute(0, 2, 1, 3)[:, :, :U].reshape(2, 4 * U, I)
        # recurrent weights: dW_hh = sum_t da_t h_{t-1}^T, "t-1" in the direction the recurrence ran
        dGv = dG.view(rows, T, 2, Up * 4)
        Hv = H.view(rows, T, 2, Up)
        if T > 1:
            gf = _mm_f32(dGv[:, 1:, 0].reshape(-1, 4 * Up).t(), Hv[:, :-1, 0].reshape(-1, Up))
            gr = _mm_f32(dGv[:, :-1, 1].reshape(-1, 4 * Up).t(), Hv[:, 1:, 1].reshape(-1, Up))
        else:
            gf = gr = torch.zeros((4 * Up, Up), dtype=torch.float32, device=dev)
        gwh = torch.stack([gf, gr]).view(2, Up, 4, Up).permute(0, 2, 1, 3)[:, :, :U, :U].reshape(2, 4 * U, U)
        dx = None
        if ctx.needs_input_grad[0]:
            wit, ld_wit = pack.w_ih_t()
            dxf = torch.empty((M, I), dtype=torch.float32, device=dev)
            ops.gemm(dG2, 8 * Up, wit, ld_wit, M, I, 8 * Up, dxf, mode=ops.EPI_F32, ldo=I, b_mod=1)
            dx = dxf.view(rows, T, I)
        return (dx, gwi[0], gwi[1], gwh[0], gwh[1], gb[0], gb[1], gb[0].clone(), gb[1].clone(), g_w_proj, g_b_proj, None,
                None)


def rnnp_layer(lstm: torch.nn.LSTM, linear: torch.nn.Linear, pack, x: torch.Tensor, act_tanh: bool) -> torch.Tensor:
    """Differentiable projected BLSTM layer on (rows, T, I)."""
    return RNNPLayerFn.apply(x, lstm.weight_ih_l0, lstm.weight_ih_l0_reverse, lstm.weight_hh_l0, lstm.weight_hh_l0_reverse,
                             lstm.bias_ih_l0, lstm.bias_ih_l0_reverse, lstm.bias_hh_l0, lstm.bias_hh_l0_reverse,
                             linear.weight, linear.bias, pack, act_tanh)


class ISTFTFn(torch.autograd.Function):
    """time = fe.istft(X, num_samples) with the adjoint as backward: the gradient wrt X is the STFT of the output
    gradient taken with the SYNTHESIS window, scaled by c_k / size (c_k = 2 for the interior bins, 1 for DC and
    Nyquist) -- irfft, window, overlap-add and trimming are all linear."""

    @staticmethod
    def forward(ctx, X, fe, num_samples):
        ctx.fe, ctx.frames = fe, X.shape[-2]
        return fe.istft(X.detach(), num_samples=num_samples)

    @staticmethod
    def backward(ctx, g):
        fe = ctx.fe
        total = (ctx.frames - 1) * fe.shift + fe.window_length - (2 * (fe.window_length - fe.shift) if fe.fading else 0)
        g = g.contiguous().float()
        if g.shape[-1] < total:  # samples cut off by num_samples carry no gradient
            g = torch.nn.functional.pad(g, (0, total - g.shape[-1]))
        dX = fe.stft(g, _window="synwin")[..., :ctx.frames, :]
        scale = torch.full((fe.frequencies,), 2.0 / fe.size, dtype=torch.float32, device=g.device)
        scale[0] = scale[-1] = 1.0 / fe.size
        return dX * scale, None, None
