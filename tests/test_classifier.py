"""The consumer classifier: layer names match the reference (state dicts load) and the fused inference
form is exact up to rounding.  CPU."""

import torch

import audiopure_b200 as ap
from oracle import resnext as o_resnext


def _randomised_bn(clf, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in clf.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return clf


def test_state_dict_layout_and_forward_match_oracle(golden):
    sd = o_resnext.make_state_dict(4321)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    assert list(clf.state_dict().keys()) == list(sd.keys())
    clf.load_state_dict(sd)
    clf.eval()
    spec = torch.from_numpy(golden("mel.npz")["logmel"])
    with torch.no_grad():
        got = clf(spec)
    want = torch.from_numpy(golden("resnext.npz")["logits"])  # the reference module's own output
    assert float((got - want).norm() / want.norm()) < 1e-5


def test_fused_inference_form_is_exact_in_fp32():
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(o_resnext.make_state_dict(4321))
    clf = _randomised_bn(clf.eval())
    fused = ap.FusedResNeXt(clf, dtype=torch.float32)
    x = torch.randn(4, 1, 32, 32, generator=torch.Generator().manual_seed(1)) * 20 - 30
    with torch.no_grad():
        a, b = clf(x), fused(x)
    assert float((a - b).norm() / a.norm()) < 1e-5
    assert sum(1 for m in fused.modules() if isinstance(m, torch.nn.BatchNorm2d)) == 0


def test_fused_form_is_inference_only():
    """FusedResNeXt has no backward; an input that requires grad must fail loudly, not lose its gradient."""
    import pytest

    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(o_resnext.make_state_dict(4321))
    fused = ap.FusedResNeXt(clf.eval(), dtype=torch.float32)
    x = torch.zeros(1, 1, 32, 32, requires_grad=True)
    with pytest.raises(RuntimeError, match="inference-only"):
        fused(x)
    with torch.no_grad():
        assert fused(x).shape == (1, 10)
