"""CPU tests of the host-side mirror of the reference interface: config protocol, factory
remapping, module tree / state_dict contract, error behaviour without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

REF = "/root/reference/tssep/exp"


def test_get_config_layout_matches_reference_doctest():
    """Dict layout pinned by tssep/train/net.py:345-365."""
    from tssep_b200.net import MaskEstimator_v2

    cfg = MaskEstimator_v2.get_config({"combination": "cat"})
    assert list(cfg.items()) == [
        ("factory", "tssep_b200.net.MaskEstimator_v2"), ("idim", 80), ("odim", None), ("layers", 3), ("units", 300),
        ("projs", 320), ("dropout", 0), ("nmask", 1), ("pre_net", "RNNP"), ("aux_net", None),
        ("aux_net_output_size", 100), ("combination", "cat"), ("ts_vad", False), ("output_resolution", "tf"),
        ("random_speaker_order", True), ("num_averaged_permutations", 1), ("input_normalizer", None),
        ("aux_normalizer", None), ("explicit_vad", False)]


def test_model_default_config_and_param_count():
    """tssep/train/model.py:74-115 (defaults) and :553-554 (114038 parameters)."""
    from tssep_b200.model import Model

    cfg = Model.get_config()
    assert cfg["fe"] == {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT", "size": 1024, "shift": 256,
                         "window_length": 1024, "pad": True, "fading": True, "output_size": 513, "window": "hann",
                         "statistics_axis": "tf"}
    assert cfg["mask_estimator"]["idim"] == 513 and cfg["mask_estimator"]["odim"] == 513
    assert cfg["mask_estimator"]["nmask"] == 1 and cfg["loss"]["target"] == "speaker_reverberation_early_ch0"
    torch.manual_seed(0)
    model = Model.new({"mask_estimator": {"units": 10, "projs": 12}})
    assert sum(p.numel() for p in model.parameters()) == 114038


def test_same_random_init_and_state_dict_keys_as_oracle():
    """Modules are created in the reference's order, so torch.manual_seed reproduces its init."""
    from tssep_b200.net import MaskEstimator_v2

    kw = dict(idim=553, odim=513, units=8, projs=6, combination="mul", ts_vad=8, aux_net_output_size=513,
              num_averaged_permutations=2)
    torch.manual_seed(3)
    ref = O.OracleMaskEstimator(**kw)
    torch.manual_seed(3)
    me = MaskEstimator_v2.new(kw)
    a, b = ref.state_dict(), me.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_repr_matches_reference_layout():
    """tssep/train/net.py:403-440 (mul, ts_vad=4, idim=513)."""
    from tssep_b200.net import MaskEstimator_v2

    r = repr(MaskEstimator_v2.new({"combination": "mul", "ts_vad": 4, "idim": 513}))
    for line in ["combination='mul',", "(0): LSTM(513, 300, batch_first=True, bidirectional=True)",
                 "(1): Linear(in_features=600, out_features=513, bias=True)", "(dropout0): Dropout(p=0, inplace=False)",
                 "(activation1): Tanh()", "(rearrange1):", "(0): LSTM(1280, 300, batch_first=True, bidirectional=True)",
                 "(linear2): Linear(in_features=320, out_features=2052, bias=True)", "(rearrange2):",
                 "(final_activation): Sigmoid()"]:
        assert line in r, line


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("variant,resolution", [("init_cfg_tssep.yaml", "tf"), ("init_cfg_tsvad.yaml", "t")])
def test_reference_yaml_loads_with_only_factories_swapped(variant, resolution):
    import yaml

    from tssep_b200.configurable import import_class, remap_factories

    def merge(a, b):
        for k, v in b.items():
            if isinstance(v, dict) and isinstance(a.get(k), dict):
                merge(a[k], v)
            else:
                a[k] = v

    cfg = yaml.safe_load(open(f"{REF}/init_cfg_common.yaml"))["eg"]["trainer"]["model"]
    merge(cfg, yaml.safe_load(open(f"{REF}/{variant}"))["eg"]["trainer"]["model"])
    cfg = remap_factories(cfg)
    cls = import_class(cfg.pop("factory"))
    model = cls.from_config(cls.get_config(cfg))
    me = model.mask_estimator
    assert me.output_resolution == resolution and me.ts_vad == 8 and me.num_averaged_permutations == 2
    assert model.fe.output_size == 553 and model.fe.frequencies == 513
    assert me.pre_net.net[0].input_size == 553 and me.post_net.birnn2.net[0].input_size == 42 * 8
    keys = list(model.state_dict().keys())
    assert "fe.fe1.dct_mat" in keys and "mask_estimator.post_net.linear2.weight" in keys
    assert model.mask_estimator.post_net.linear2.out_features == (4104 if resolution == "tf" else 8)


def test_vad2sep_broadcast_relies_on_speaker_major_head_rows():
    """InitCheckPointVAD2Sep (tssep/train/init_ckpt.py:62-83): repeat_interleave of the 't' head gives
    a 'tf' head whose logits are the 't' logits broadcast over frequency."""
    kw = dict(idim=40, odim=33, units=4, projs=5, combination="mul", ts_vad=3, aux_net_output_size=33,
              random_speaker_order=False)
    torch.manual_seed(0)
    vad = O.OracleMaskEstimator(output_resolution="t", **kw)
    sep = O.OracleMaskEstimator(output_resolution="tf", **kw)
    sd = vad.state_dict()
    for k in ["post_net.linear2.weight", "post_net.linear2.bias"]:
        sd[k] = torch.repeat_interleave(sd[k], 33, dim=0)
    sep.load_state_dict(sd)
    xs, aux = torch.rand(12, 40), [torch.rand(33) for _ in range(3)]
    with torch.no_grad():
        assert (vad(xs, aux).logit - sep(xs, aux).logit).abs().max().item() < 1e-6


def test_dummy_reader_matches_oracle_generator():
    from tssep_b200.data import DummyReader

    ex = DummyReader(aux_size=513).get_example(3)
    want = O.dummy_example(3, aux_size=513)
    assert np.array_equal(ex["audio_data"]["observation"], want["observation"])
    assert np.array_equal(ex["auxInput"], want["auxInput"])
    assert np.array_equal(ex["audio_data"]["speaker_reverberation_early_ch0"], want["speaker_reverberation_early_ch0"])
    assert ex["audio_data"]["observation"].shape == (1, 80000) and ex["auxInput"].shape == (8, 513)


def test_frame_geometry_helpers():
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    fe = Log1pMaxNormAbsSTFT()
    assert repr(fe) == ("Log1pMaxNormAbsSTFT(size=1024, shift=256, window_length=1024, pad=True, fading=True, "
                        "output_size=513, window='blackman', statistics_axis='tf')")
    assert [fe.num_frames(n) for n in (80000, 10000, 9600000)] == [316, 43, 37503]
    s = np.arange(0, 4000, 13)
    assert (fe.sample_index_to_frame_index(s) == O.sample_to_frame_index(s, 1024, 256)).all()


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing silently falls back to PyTorch."""
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT
    from tssep_b200.net import MaskEstimator_v2
    from tssep_b200.rnnp import RNNP_packed

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Log1pMaxNormAbsSTFT().stft(torch.zeros(2000))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RNNP_packed(8, 1, 4, 4, 0)(torch.zeros(5, 8))
    me = MaskEstimator_v2.new({"idim": 16, "units": 4, "projs": 4})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        me(torch.zeros(5, 16), [torch.zeros(100)] * 3)


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tssep_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_unsupported_options_raise():
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT
    from tssep_b200.rnnp import RNNP_packed

    with pytest.raises(NotImplementedError):
        RNNP_packed(8, 1, 4, 4, 0, typ="bgru")
    with pytest.raises(NotImplementedError):
        Log1pMaxNormAbsSTFT(statistics_axis="t")._feature_parts()


def test_operand_leading_dimension_is_128_byte_aligned_for_long_rows():
    """ops.operand_ld: bf16 GEMM operand rows of 128 bytes or more start on 128-byte boundaries (64 elements),
    short rows only need the 16-byte TMA alignment (8 elements)."""
    from tssep_b200 import ops

    assert ops.operand_ld(513) == 576 and ops.operand_ld(553) == 576 and ops.operand_ld(320) == 320
    assert ops.operand_ld(2560) == 2560 and ops.operand_ld(640) == 640 and ops.operand_ld(64) == 64
    assert ops.operand_ld(33) == 40 and ops.operand_ld(12) == 16 and ops.operand_ld(3) == 8
    for cols in range(1, 700):
        ld = ops.operand_ld(cols)
        assert ld >= cols and ld % 8 == 0 and (cols < 64 or ld % 64 == 0) and ld - cols < 64


def test_recurrence_kernel_selection(monkeypatch):
    """TSSEP_LSTM_KERNEL: the tensor-memory kernel for every row count whenever the input projections are bf16 (the
    default), the register kernel for f32 projections or on request."""
    import torch

    from tssep_b200 import rnnp

    monkeypatch.delenv("TSSEP_LSTM_KERNEL", raising=False)
    assert rnnp.rec_kernel(torch.bfloat16, 304) == "ts" and rnnp.rec_kernel(torch.bfloat16, 16) == "ts"
    assert rnnp.rec_kernel(torch.float32, 304) == "regs"
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", "regs")
    assert rnnp.rec_kernel(torch.bfloat16, 304) == "regs"
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", "ts")
    assert rnnp.rec_kernel(torch.bfloat16, 48) == "ts"


def test_mask_estimator_can_be_deep_copied_and_pickled():
    """The kernel-side caches of the module (packed weights, pinned staging ring) are derived state: copies and
    pickles of the module carry the parameters only."""
    import copy
    import io

    import torch

    from tssep_b200.net import MaskEstimator_v2

    me = MaskEstimator_v2.new(dict(idim=80, odim=33, units=4, projs=6, combination="cat", aux_net_output_size=10,
                                   ts_vad=False, num_averaged_permutations=1))
    me2 = copy.deepcopy(me)
    assert me2._staging is not me._staging
    assert all(torch.equal(a, b) for a, b in zip(me.state_dict().values(), me2.state_dict().values()))
    buf = io.BytesIO()
    torch.save(me, buf)
    buf.seek(0)
    me3 = torch.load(buf, weights_only=False)
    assert list(me3.state_dict().keys()) == list(me.state_dict().keys())


def test_stream_handle_carries_its_device():
    """_lib.StreamHandle is the integer cudaStream_t plus the device it belongs to; _lib.call uses it as the device
    guard of the launch (advisor finding of round 1: kernels used to launch on the current device with another
    device's stream)."""
    import ctypes

    from tssep_b200 import _lib

    h = _lib.StreamHandle(0x1234, 3)
    assert int(h) == 0x1234 and h.device_index == 3 and isinstance(h, int)
    assert ctypes.c_void_p(h).value == 0x1234  # converts like a plain integer at the ctypes boundary
    args = (1, 2.0, None, h)
    assert next((a.device_index for a in args if isinstance(a, _lib.StreamHandle)), -1) == 3


def test_instance_norm_modules_mirror_the_reference_signatures():
    from tssep_b200.configurable import FACTORY_ALIASES
    from tssep_b200.net import InstanceNorm, InstanceNorm_v2

    assert repr(InstanceNorm(dim=-1)) == "InstanceNorm(dim=-1, unbiased=False)"          # net.py:257-258
    assert repr(InstanceNorm_v2(-1, -2)) == "InstanceNorm_v2(mean_dim=-1, norm_dim=-2)"
    assert FACTORY_ALIASES["tssep.train.net.InstanceNorm_v2"] == "tssep_b200.net.InstanceNorm_v2"


def test_row_pitch_view():
    """The mask consumers read the head GEMM's padded buffers in place; anything else is made contiguous."""
    import torch

    from tssep_b200.ops import row_pitch_view

    buf = torch.zeros((2, 3, 1, 7, 520))
    v = buf[..., :513]
    t, p = row_pitch_view(v)
    assert p == 520 and t.data_ptr() == buf.data_ptr() and t.shape == v.shape
    t, p = row_pitch_view(v[0])                      # un-batched view
    assert p == 520 and t.data_ptr() == buf.data_ptr()
    t, p = row_pitch_view(v[1, 1:])                  # sliced over a dense leading axis
    assert p == 520 and t.data_ptr() == v[1, 1:].data_ptr()
    c = torch.zeros((2, 3, 1, 7, 513))
    t, p = row_pitch_view(c)
    assert p == 513 and t.data_ptr() == c.data_ptr()
    t, p = row_pitch_view(buf[:, :, :, ::2, :513])   # rows not equidistant with the leading axes -> copy
    assert p == 513 and t.is_contiguous()
    t, p = row_pitch_view(v.double())                # wrong dtype -> float32 copy
    assert p == 513 and t.dtype == torch.float32 and t.is_contiguous()
    t, p = row_pitch_view(buf[..., 1:514])           # offset start is fine as long as rows stay equidistant
    assert p == 520 and t.data_ptr() == buf[..., 1:514].data_ptr()


def test_choose_ts_shape_respects_the_cta_budget():
    """Two steps in flight: every recurrence launch gets half of the 148 SMs."""
    from tssep_b200.rnnp import choose_ts_shape

    cap = {(8, 1, 1): 56, (8, 2, 1): 104, (16, 2, 2): 208, (32, 2, 2): 416, (64, 2, 2): 832}
    capacity = lambda Up, rpc, tiles, subs: cap.get((rpc, tiles, subs), 0)
    # (rows per cluster, tiles, subs)
    assert choose_ts_shape(8, 304, 74, capacity) == (8, 1, 1)        # 20 CTAs, the latency shape
    assert choose_ts_shape(16, 304, 74, capacity) == (8, 1, 1)       # 40 CTAs
    assert choose_ts_shape(64, 304, 74, capacity) == (16, 2, 2)      # (8, 2, 1) would need 80 CTAs
    assert choose_ts_shape(64, 304, 148, capacity) == (8, 2, 1)
    assert choose_ts_shape(128, 304, 74, capacity) == (32, 2, 2)     # 4 clusters of 5 per direction = 40 CTAs
    assert choose_ts_shape(416, 304, 74, capacity) == (64, 2, 2)     # 7 clusters x 5 x 2 = 70 CTAs
    assert choose_ts_shape(832, 304, 74, capacity) == (0, 0, 0)      # nothing fits half the device: library's choice


def test_classic_bf_config_protocol():
    """``ClassicBF_np.get_config()`` / ``new()`` as the reference's doctest shows them (enhancer.py:386-398), and
    reference factory strings inside a config resolve to this package's classes."""
    from tssep_b200.enhancer import WPE, ClassicBF_np
    from tssep_b200.enhancer_distortion_mask import SumCrossTalker

    cfg = ClassicBF_np.get_config()
    assert cfg == {"factory": "tssep_b200.enhancer.ClassicBF_np", "bf": "mvdr_souden", "masking": False, "masking_eps": 0,
                   "distortion_mask": {"factory": "tssep_b200.enhancer_distortion_mask.SumCrossTalker", "eps": 0.0001},
                   "pre_wpe": None, "segment_wpe": None, "mask_power": 1}
    enh = ClassicBF_np.new()
    assert isinstance(enh.distortion_mask, SumCrossTalker) and enh.distortion_mask.eps == 0.0001
    enh = ClassicBF_np.new({"pre_wpe": {"factory": "tssep.train.enhancer.WPE", "taps": 5}, "bf": "ch0",
                            "distortion_mask": {"eps": 0.01}})
    assert isinstance(enh.pre_wpe, WPE) and enh.pre_wpe.taps == 5 and enh.pre_wpe.delay == 2
    assert enh.bf == "ch0" and enh.distortion_mask.eps == 0.01
