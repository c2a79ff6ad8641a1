"""GPU parity: tcgen05 GEMM (and the SIMT bisecting kernel) vs torch fp32 matmul on the
same bf16-rounded operands.  Accumulation is fp32 in both, so the tolerance only covers
summation order: 1e-3 relative to the largest output magnitude."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


def _run(cuda, M, N, K, impl, batch=1, a_div=1, shared_b=False, act=0, out_bf16=False, alpha=1.0):
    from tssep_b200 import ops

    n_a = (batch + a_div - 1) // a_div
    A = _rand((n_a, M, K), 1).to(cuda)
    B = _rand((1 if shared_b else batch, N, K), 2).to(cuda) / np.sqrt(K)
    bias = _rand((batch, N), 3).to(cuda)
    lda, ldb = ops.round_up(K, 8), ops.round_up(K, 8)
    Ab = ops.cast_bf16(A.reshape(-1, K), lda)
    Bb = ops.cast_bf16(B.reshape(-1, K), ldb)
    ldo = ops.round_up(N, 8) if out_bf16 else N
    out = torch.zeros((batch, M, ldo), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=cuda)
    ops.gemm(Ab, lda, Bb, ldb, M, N, K, out, mode=ops.EPI_BF16 if out_bf16 else ops.EPI_F32, ldo=ldo, batch=batch,
             a_stride=M * lda, a_div=a_div, b_stride=0 if shared_b else N * ldb, bias=bias, bias_stride=N,
             alpha=alpha, act=act, out_stride=M * ldo, impl=impl)
    torch.cuda.synchronize()
    Af = Ab[:, :K].float().reshape(n_a, M, K)
    Bf = Bb[:, :K].float().reshape(-1, N, K)
    want = torch.stack([alpha * Af[z // a_div] @ Bf[0 if shared_b else z].T + bias[z] for z in range(batch)])
    if act:
        want = torch.tanh(want)
    got = out[..., :N].float()
    tol = (1e-3 if not out_bf16 else 1e-2) * max(1.0, want.abs().max().item())
    err = (got - want).abs().max().item()
    assert err < tol, (err, tol)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 100, 70), (1000, 2432, 553), (129, 513, 608), (37, 8, 42),
                                   (4096, 4104, 640)])
def test_gemm_plain(cuda, impl, M, N, K):
    _run(cuda, M, N, K, impl)


@pytest.mark.parametrize("impl", [1, 0])
def test_gemm_batched_shared_a(cuda, impl):
    _run(cuda, 316, 384, 513, impl, batch=16, a_div=8)


@pytest.mark.parametrize("impl", [1, 0])
def test_gemm_shared_b_tanh_bf16(cuda, impl):
    _run(cuda, 500, 320, 608, impl, batch=3, shared_b=True, act=1, out_bf16=True, alpha=0.5)


@pytest.mark.parametrize("impl", [1, 0])
def test_gemm_head_scatter(cuda, impl):
    from tssep_b200 import ops

    Z, K_spk, F, T, P = 2, 4, 33, 150, 24
    A = _rand((Z, T, P), 1).to(cuda)
    W = _rand((K_spk * F, P), 2).to(cuda) / np.sqrt(P)
    b = _rand((K_spk * F,), 3).to(cuda)
    ld = ops.round_up(P, 8)
    Ab, Wb = ops.cast_bf16(A.reshape(-1, P), ld), ops.cast_bf16(W, ld)
    rng = np.random.RandomState(0)
    perm = np.stack([rng.permutation(K_spk) for _ in range(Z)])
    planes = (np.arange(Z)[:, None] * K_spk + perm).astype(np.int32)
    logit = torch.zeros((Z, K_spk, 1, T, F), device=cuda)
    mask = torch.zeros_like(logit)
    ops.gemm(Ab, ld, Wb, ld, T, K_spk * F, P, logit, mode=ops.EPI_HEAD, batch=Z, a_stride=T * ld, b_stride=0, b_mod=1,
             bias=b, alpha=0.5, mask=mask, plane_map=torch.tensor(planes.reshape(-1), device=cuda), n_blocks=K_spk,
             row_len=F, impl=impl)
    torch.cuda.synchronize()
    full = 0.5 * Ab[:, :P].float().reshape(Z, T, P) @ Wb[:, :P].float().T + b  # (Z,T,K*F)
    full = full.reshape(Z, T, K_spk, F).permute(0, 2, 1, 3)  # slot order
    want = torch.zeros_like(logit)
    for z in range(Z):
        for q in range(K_spk):
            want[z, perm[z, q], 0] = full[z, q]
    assert (logit - want).abs().max().item() < 2e-3
    assert (mask - torch.sigmoid(want)).abs().max().item() < 1e-3
