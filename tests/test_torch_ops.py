"""The torch custom-op layer (tssep_b200/torch_ops.py): one ``torch.ops.tssep_b200`` operator per compute entry point of
the C ABI, CUDA-only kernels, fake implementations so that shapes propagate and Dynamo traces through them."""
import re

import pytest
import torch

from tests.test_lib_exports import declared_symbols

HOST_QUERIES = {"tssep_last_error", "tssep_abi_version", "tssep_device_info", "tssep_blstm_recurrence_ts_capacity",
                "tssep_wpe_workspace_bytes"}


class RnnpLayer(torch.nn.Module):
    """x . W_ih^T -> recurrence -> projection: the three launches of one RNNP layer, written against tssep_b200.ops."""

    def __init__(self, I, U, P):
        super().__init__()
        from tssep_b200 import ops

        self.I, self.Up, self.P = I, ops.round_up(U, 16), P

    def forward(self, x, w_ih, bias, wimg, w_proj, b_proj):
        from tssep_b200 import ops

        rows, T, _ = x.shape
        Up = self.Up
        xb = ops.cast_bf16(x.reshape(rows * T, self.I))
        ld = ops.operand_ld(self.I)
        G = torch.empty((rows * T, 8 * Up), dtype=torch.bfloat16, device=x.device)
        ops.gemm(xb, ld, w_ih, ld, rows * T, 8 * Up, self.I, G, mode=ops.EPI_BF16, ldo=8 * Up, bias=bias)
        H = ops.blstm_recurrence_ts(G, wimg, rows, T, Up, fast_math=True)
        out = torch.empty((rows * T, self.P), dtype=torch.float32, device=x.device)
        ops.gemm(H, 2 * Up, w_proj, 2 * Up, rows * T, self.P, 2 * Up, out, mode=ops.EPI_F32, ldo=self.P, bias=b_proj,
                 b_mod=1)
        return out.reshape(rows, T, self.P)


def test_every_compute_entry_point_has_an_operator():
    from tssep_b200 import torch_ops

    assert sorted(set(declared_symbols()) - HOST_QUERIES) == torch_ops.WRAPPED_SYMBOLS
    for name in torch_ops._OPS:
        assert hasattr(torch.ops.tssep_b200, name)


def test_operator_schemas_mark_outputs_as_mutated():
    from tssep_b200 import torch_ops

    for name, (schema, _) in torch_ops._OPS.items():
        s = getattr(torch.ops.tssep_b200, name).default._schema
        assert any(a.alias_info is not None and a.alias_info.is_write for a in s.arguments), name
        assert len(s.returns) == 0, name
        assert re.search(r"Tensor\([a-z]!\)", schema), name


def test_cpu_tensors_have_no_kernel():
    """CUDA dispatch key only: a CPU tensor cannot reach the library (no CPU fallback anywhere on the product path)."""
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.tssep_b200.cast_bf16(torch.zeros(4, 4), 4, 4, 4, torch.zeros((4, 8), dtype=torch.bfloat16), 8)
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.tssep_b200.activity(torch.zeros(1, 3, 5), 1, 3, 5, 0, torch.zeros(1, 3))


def test_dynamo_traces_a_layer_without_graph_breaks():
    """torch.export(strict=True) = Dynamo with fullgraph semantics, on FAKE cuda tensors (no GPU, nothing executes): the
    graph of one RNNP layer holds the four operators and no graph break."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    m = RnnpLayer(64, 40, 42)
    Up = 48
    with FakeTensorMode():
        dev = "cuda"
        args = (torch.empty((3, 50, 64), device=dev), torch.empty((8 * Up, 64), dtype=torch.bfloat16, device=dev),
                torch.empty(8 * Up, device=dev), torch.empty(2 * 1 * 2 * 3 * 128 * 8, dtype=torch.int32, device=dev),
                torch.empty((42, 2 * Up), dtype=torch.bfloat16, device=dev), torch.empty(42, device=dev))
        ep = torch.export.export(m, args, strict=True)
    targets = [str(n.target) for n in ep.graph.nodes if n.op == "call_function" and "tssep_b200" in str(n.target)]
    assert targets == ["tssep_b200.cast_bf16.default", "tssep_b200.gemm.default",
                       "tssep_b200.blstm_recurrence_ts.default", "tssep_b200.gemm.default"], targets
    out = [n for n in ep.graph.nodes if n.op == "output"][0]
    assert tuple(out.args[0][-1].meta["val"].shape) == (3, 50, 42)


@pytest.mark.gpu
def test_compiled_layer_equals_eager(cuda):
    """torch.compile(fullgraph=True) of the same layer on the GPU: identical output to the eager call."""
    from tssep_b200 import ops

    torch.manual_seed(0)
    I, U, P = 64, 40, 42
    Up = ops.round_up(U, 16)
    lstm = torch.nn.LSTM(I, U, bidirectional=True, batch_first=True).to(cuda)
    w = torch.zeros((2, 4, Up, I), device=cuda)
    w[0, :, :U] = lstm.weight_ih_l0.detach().view(4, U, I)
    w[1, :, :U] = lstm.weight_ih_l0_reverse.detach().view(4, U, I)
    w_ih = ops.cast_bf16(w.view(8 * Up, I))
    bias = torch.zeros(8 * Up, device=cuda)
    wimg = ops.pack_whh_ts(lstm.weight_hh_l0.detach().contiguous(), lstm.weight_hh_l0_reverse.detach().contiguous(), U, Up)
    w_proj = ops.cast_bf16(torch.randn((P, 2 * Up), device=cuda) * 0.1, 2 * Up)
    b_proj = torch.randn(P, device=cuda)
    x = torch.randn((3, 50, I), device=cuda)
    m = RnnpLayer(I, U, P)
    want = m(x, w_ih, bias, wimg, w_proj, b_proj)
    got = torch.compile(m, fullgraph=True, backend="eager")(x, w_ih, bias, wimg, w_proj, b_proj)
    assert torch.equal(got, want)
