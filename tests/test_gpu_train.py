"""GPU parity of the training step (BASELINE config 5): forward AND backward of the projected BLSTM layers on this
package's kernels (tensor-memory recurrence that stores its gates, BPTT kernel csrc/lstm_bwd.cu, tcgen05 GEMMs for the
data gradients) against torch autograd through the oracle (f32, CPU).

bf16 operands with f32 accumulation: gradients carry the rounding of every bf16 operand of the chain; the stated bar is
a relative L2 error <= 5e-2 and a cosine similarity >= 0.995 per parameter (measured values are printed)."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL, COS_TOL = 5e-2, 0.995


def _cmp(name, got, want, stats):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (name, got.shape, want.shape)
    rel = ((got - want).norm() / (want.norm() + 1e-30)).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    stats.append((name, rel, cos))
    return rel, cos


def _report(stats, title):
    worst = max(stats, key=lambda s: s[1])
    print(f"{title}: {len(stats)} tensors, worst rel L2 {worst[1]:.3e} ({worst[0]}), min cosine {min(s[2] for s in stats):.6f}")
    for name, rel, cos in stats:
        assert rel <= REL_TOL and cos >= COS_TOL, (name, rel, cos)


@pytest.mark.parametrize("idim,units,hdim,shape", [
    (64, 40, 42, (3, 50, 64)),        # one CTA
    (80, 128, 64, (9, 70, 80)),       # cluster of 2, two row groups
    (160, 300, 320, (5, 120, 160)),   # full-size units: cluster of 5, three m tiles in the BPTT kernel
    (160, 300, 320, (12, 90, 160)),
    (72, 10, 12, (2, 40, 72)),        # Up = 16: one k-step
])
def test_rnnp_layer_gradients_match_torch(cuda, idim, units, hdim, shape):
    from tssep_b200.rnnp import RNNP_packed

    torch.manual_seed(0)
    ref = O._RNNP(idim, units, hdim)
    mine = RNNP_packed(idim, 1, units, hdim, 0)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.to(cuda).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g)
    w = torch.randn((*shape[:-1], hdim), generator=g)
    xr = x.clone().requires_grad_(True)
    (ref(xr) * w).sum().backward()
    xm = x.to(cuda).requires_grad_(True)
    ym = mine(xm)
    assert ym.requires_grad
    (ym * w.to(cuda)).sum().backward()
    stats = []
    _cmp("y", ym, ref(x), stats)
    _cmp("dx", xm.grad, xr.grad, stats)
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        _cmp(n, p.grad, q.grad, stats)
    _report(stats, f"RNNP layer U={units} rows={int(np.prod(shape[:-2]))} T={shape[-2]}")


def test_istft_adjoint(cuda):
    """tssep_b200.autograd.ISTFTFn: gradient of sum(w * istft(X)) wrt X vs torch autograd through the oracle iSTFT."""
    from tssep_b200.autograd import ISTFTFn
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    for size, shift, n in ((1024, 256, 5000), (256, 64, 3333)):
        fe = Log1pMaxNormAbsSTFT(size=size, shift=shift, window="hann")
        T = fe.num_frames(n)
        g = torch.Generator().manual_seed(0)
        X = torch.view_as_complex(torch.randn((2, T, size // 2 + 1, 2), generator=g))
        w = torch.randn((2, n), generator=g)
        Xr = X.clone().requires_grad_(True)
        (O.istft(Xr, size=size, shift=shift, window="hann", num_samples=n) * w).sum().backward()
        Xm = X.to(cuda).requires_grad_(True)
        (ISTFTFn.apply(Xm, fe, n) * w.to(cuda)).sum().backward()
        d = torch.view_as_real(Xm.grad.cpu() - Xr.grad)
        # the imaginary parts of DC / Nyquist do not enter irfft: torch reports their gradient as zero as well
        rel = (d.norm() / torch.view_as_real(Xr.grad).norm()).item()
        print(f"iSTFT adjoint {size}/{shift}: rel L2 {rel:.3e}")
        assert rel < 1e-5, rel


@pytest.mark.parametrize("kw", [
    dict(combination="mul", ts_vad=8, num_averaged_permutations=2, aux_net_output_size=513, output_resolution="tf"),
    dict(combination="cat", ts_vad=False, num_averaged_permutations=1, aux_net_output_size=100, output_resolution="tf"),
    dict(combination="mul", ts_vad=8, num_averaged_permutations=2, aux_net_output_size=513, output_resolution="t"),
])
def test_training_step_gradients_match_oracle(cuda, kw):
    """One training step of the toy TS-SEP / TS-VAD model: Model.forward in train mode -> LogMAE (tf) or VADSigmoidBCE
    (t) -> backward; the loss and the gradient of all 42 parameters (net.py:735-776) against oracle autograd."""
    from tests.util import make_pair
    from tssep_b200.data import DummyReader
    from tssep_b200.enhancer import Masking
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures
    from tssep_b200.loss import LogMAE, VADSigmoidBCE
    from tssep_b200.model import Model

    base = dict(idim=553, odim=513, units=40, projs=42)
    base.update(kw)
    ref, me = make_pair(base, device=cuda)
    ref.train()
    fe = ConcaternatedSTFTFeatures.new({
        "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
        "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
        "size": 1024, "shift": 256, "window": "hann"})
    model = Model(fe=fe, reader=DummyReader(), mask_estimator=me, enhancer=Masking(), loss=LogMAE()).to(cuda).train()
    A = kw["aux_net_output_size"]
    n = 16000 * 2
    exs = [O.dummy_example(s, aux_size=A, num_samples=n) for s in range(2)]
    obs = torch.tensor(np.stack([e["observation"] for e in exs]))          # (2, 1, n)
    aux = torch.tensor(np.stack([e["auxInput"] for e in exs]))
    tgt = torch.tensor(np.stack([e["speaker_reverberation_early_ch0"] for e in exs]))
    vad_frames = torch.tensor(np.stack([O.stft_vad(e["vad"], 1024, 256, True) for e in exs]).astype(np.float32))
    tf = kw["output_resolution"] == "tf"

    # oracle: same path, torch autograd on the CPU
    np.random.seed(4)
    X = O.stft(obs, size=1024, shift=256, window="hann")
    inp = O.concat_feature(X[:, 0], O.MFCCTables()).float()
    out = ref(inp, [[a for a in item] for item in aux])
    if tf:
        est = O.masking(out.mask, X, 0)
        time = O.istft(est, size=1024, shift=256, window="hann", num_samples=n)
        loss_ref = O.log_mae(time, tgt).sum()
    else:
        loss_ref = VADSigmoidBCE()(torch.squeeze(out.logit, dim=-3), vad_frames).sum()
    loss_ref.backward()

    np.random.seed(4)
    ex = {"observation": obs.to(cuda), "auxInput": aux.to(cuda), "reference_channel": 0}
    got = model(ex)
    if tf:
        loss = model.loss(got.time_estimate, tgt.to(cuda)).sum()
    else:
        loss = VADSigmoidBCE()(torch.squeeze(got.logit, dim=-3), vad_frames.to(cuda)).sum()
    loss.backward()
    print(f"loss {loss.item():.6f} vs oracle {loss_ref.item():.6f}")
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * max(1.0, abs(loss_ref.item()))
    stats = []
    _cmp("mask", got.mask, out.mask, stats)
    names = [n_ for n_, _ in me.named_parameters()]
    assert len(names) == 42
    ref_grads = dict(ref.named_parameters())
    for n_, p in me.named_parameters():
        assert p.grad is not None, n_
        _cmp(n_, p.grad, ref_grads[n_].grad, stats)
    _report(stats, f"training step {kw}")
