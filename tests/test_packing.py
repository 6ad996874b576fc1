"""Host logic on the CPU: state-dict layout, weight packing and the kernel dataflow (emulated in torch)
against the oracle.  No GPU, no compute through the C ABI."""

import json

import pytest
import torch

import audiopure_b200 as ap
from oracle import schedule as o_schedule, wavenet as o_wavenet, weights as W
from tests.emulate import emulate_eps, emulate_eps_tf32

SMALL = dict(W.DEFAULT_WAVENET_CONFIG, num_res_layers=6, dilation_cycle=3)


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_state_dict_layout_matches_reference():
    m = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    sd = W.make_state_dict(1234)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(m.state_dict()[k].shape == v.shape for k, v in sd.items())
    m.load_state_dict(sd)  # strict
    # default init leaves the output conv at zero, like ZeroConv1d (WaveNet.py:39-44)
    fresh = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    assert float(fresh.state_dict()["final_conv.2.conv.weight"].abs().max()) == 0.0


def test_unsupported_widths_are_rejected():
    with pytest.raises(NotImplementedError):
        ap.WaveNet_Speech_Commands(res_channels=128, skip_channels=128)


def test_schedule_matches_oracle_bit_exact():
    a = ap.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)
    b = o_schedule.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert torch.equal(a[k], b[k])
    steps = torch.tensor([[0.0], [5.0], [199.0]])
    assert torch.equal(ap.calc_diffusion_step_embedding(steps, 128), o_schedule.calc_diffusion_step_embedding(steps, 128))


@pytest.mark.parametrize("t", [0, 7, 199])
def test_packed_dataflow_fp32_equals_oracle(t):
    """With no bf16 rounding the packed reformulation (permuted rows, tap-major K, folded sqrt(.5) and
    sqrt(1/N), shift folded into the previous layer's epilogue, deferred skip GEMM) IS the reference network."""
    sd = W.make_state_dict(99, SMALL)
    m = ap.WaveNet_Speech_Commands(**SMALL)
    m.load_state_dict(sd)
    packed = {k: (v.float() if torch.is_tensor(v) else v) for k, v in m.pack_weights("cpu").items()}
    # undo the bf16 cast of the GEMM operands by re-packing in fp32
    x = W.make_waveforms(2, 1000, seed=5)
    want = o_wavenet.eps_theta(sd, x, t, SMALL)
    got = emulate_eps(packed, x, t, 6, 3, quantize=False)
    # operands were rounded to bf16 at pack time: ~2^-9 relative per weight
    assert rel_l2(got, want) < 6e-3


def test_packed_dataflow_bf16_within_gate():
    """The bf16 rounding points of the kernels keep eps inside the 2e-2 gate (SURVEY section 7 numerics)."""
    sd = W.make_state_dict(99, SMALL)
    m = ap.WaveNet_Speech_Commands(**SMALL)
    m.load_state_dict(sd)
    packed = m.pack_weights("cpu")
    x = W.make_waveforms(2, 1000, seed=5)
    want = o_wavenet.eps_theta(sd, x, 7, SMALL)
    got = emulate_eps(packed, x, 7, 6, 3, quantize=True)
    assert rel_l2(got, want) < 2e-2


def test_packed_dataflow_tf32_within_gate():
    """The tf32 build (fp32 storage, operands rounded to 10 mantissa bits) keeps eps inside the 2e-3 gate."""
    sd = W.make_state_dict(99, SMALL)
    m = ap.WaveNet_Speech_Commands(**SMALL, precision="tf32")
    m.load_state_dict(sd)
    packed = m.pack_weights("cpu")
    for k in ("w1", "w2", "ws", "wf"):
        assert packed[k].dtype == torch.float32
        assert int((packed[k].view(torch.int32) & 0x1FFF).abs().max()) == 0  # already tf32: low 13 bits clear
    x = W.make_waveforms(2, 1000, seed=5)
    want = o_wavenet.eps_theta(sd, x, 7, SMALL)
    got = emulate_eps_tf32(packed, x, 7, 6, 3)
    assert rel_l2(got, want) < 2e-3


def test_round_to_tf32_is_nearest_with_ties_away():
    from audiopure_b200.wavenet import round_to_tf32
    u = 2.0 ** -10  # tf32 ulp at 1.0
    x = torch.tensor([1.0, 1.0 + 0.25 * u, 1.0 + 0.5 * u, 1.0 + 0.75 * u, -1.0 - 0.5 * u, -1.0 - 0.49 * u, 0.0])
    want = torch.tensor([1.0, 1.0, 1.0 + u, 1.0 + u, -1.0 - u, -1.0, 0.0])
    assert torch.equal(round_to_tf32(x), want)


def test_unknown_precision_is_rejected():
    with pytest.raises(NotImplementedError):
        ap.WaveNet_Speech_Commands(**SMALL, precision="fp8")


def test_packing_is_refreshed_after_load_state_dict():
    m = ap.WaveNet_Speech_Commands(**SMALL)
    m._engine = object()
    m.load_state_dict(W.make_state_dict(99, SMALL))
    assert m._engine is None


def test_product_synthetic_checkpoints_equal_oracle_ones():
    """audiopure_b200.synthetic (used by bench.py / tools) and oracle.weights (used by tests and fixtures) must
    generate the same tensors, so benchmark numbers are on the weights the parity tests cover."""
    from audiopure_b200 import synthetic as S
    from oracle import resnext as o_resnext

    a, b = S.diffwave_state_dict(1234), W.make_state_dict(1234)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    a, b = S.resnext_state_dict(4321), o_resnext.make_state_dict(4321)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(S.waveforms(3, 1000, seed=5), W.make_waveforms(3, 1000, seed=5))
    assert torch.equal(S.clips(3, 1000, seed=5), W.make_clips(3, 1000, seed=5))
    assert torch.equal(S.noise((2, 3, 4), seed=9), W.make_noise((2, 3, 4), seed=9))
