"""Deterministic synthetic inputs and random-init checkpoints for benchmarks, smoke runs and demos.

No trained DiffWave / ResNeXt checkpoint is reachable offline, so measurements use random-init weights of the
named architectures in the reference's state-dict layouts (``WaveNet_Speech_Commands.state_dict()``: 408 tensors,
weight-normed convs as ``weight_g``/``weight_v``/``bias`` -- WaveNet.py:23-34,57-73,138-162; ``CifarResNeXt``:
resnext.py:88-111) drawn from numpy ``PCG64`` streams, so a seed gives the same tensors on every machine.
The zero-initialised output conv (``ZeroConv1d``, WaveNet.py:39-44) is re-randomised, otherwise eps == 0.
"""

import os
from collections import OrderedDict

import numpy as np
import torch

DEFAULT_WAVENET_CONFIG = {
    "in_channels": 1, "res_channels": 256, "skip_channels": 256, "out_channels": 1, "num_res_layers": 36,
    "dilation_cycle": 12, "diffusion_step_embed_dim_in": 128, "diffusion_step_embed_dim_mid": 512,
    "diffusion_step_embed_dim_out": 512,
}
DEFAULT_DIFFUSION_CONFIG = {"T": 200, "beta_0": 0.0001, "beta_T": 0.02}


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def _wn_conv(rng, sd, prefix, cout, cin, k):
    bound = 1.0 / np.sqrt(cin * k)
    v = _uniform(rng, (cout, cin, k), bound)
    g = v.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1)
    g = g * torch.from_numpy(rng.uniform(0.9, 1.1, size=(cout, 1, 1)).astype(np.float32))
    sd[prefix + ".bias"] = _uniform(rng, (cout,), bound)
    sd[prefix + ".weight_g"] = g
    sd[prefix + ".weight_v"] = v


def _linear(rng, sd, prefix, cout, cin):
    bound = 1.0 / np.sqrt(cin)
    sd[prefix + ".weight"] = _uniform(rng, (cout, cin), bound)
    sd[prefix + ".bias"] = _uniform(rng, (cout,), bound)


def diffwave_state_dict(seed=1234, wavenet_config=None):
    """Random-init DiffWave checkpoint in the reference layout (torch-default magnitudes, g = ||v|| * U(0.9, 1.1))."""
    cfg = dict(DEFAULT_WAVENET_CONFIG)
    if wavenet_config:
        cfg.update(wavenet_config)
    C, S = cfg["res_channels"], cfg["skip_channels"]
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()
    _wn_conv(rng, sd, "init_conv.0.conv", C, cfg["in_channels"], 1)
    _linear(rng, sd, "residual_layer.fc_t1", cfg["diffusion_step_embed_dim_mid"], cfg["diffusion_step_embed_dim_in"])
    _linear(rng, sd, "residual_layer.fc_t2", cfg["diffusion_step_embed_dim_out"], cfg["diffusion_step_embed_dim_mid"])
    for n in range(cfg["num_res_layers"]):
        p = "residual_layer.residual_blocks.%d" % n
        _linear(rng, sd, p + ".fc_t", C, cfg["diffusion_step_embed_dim_out"])
        _wn_conv(rng, sd, p + ".dilated_conv_layer.conv", 2 * C, C, 3)
        _wn_conv(rng, sd, p + ".res_conv", C, C, 1)
        _wn_conv(rng, sd, p + ".skip_conv", S, C, 1)
    _wn_conv(rng, sd, "final_conv.0.conv", S, S, 1)
    sd["final_conv.2.conv.weight"] = torch.from_numpy(
        rng.normal(0.0, 0.05, size=(cfg["out_channels"], S, 1)).astype(np.float32))
    sd["final_conv.2.conv.bias"] = torch.from_numpy(rng.normal(0.0, 0.01, size=(cfg["out_channels"],)).astype(np.float32))
    return sd


CALIB_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "resnext_calib.npz")


def resnext_state_dict(seed=4321, nlabels=10, in_channels=1, cardinality=8, depth=29, base_width=64, widen=4,
                       calibrated=True):
    """Synthetic ResNeXt-29 8x64 checkpoint in the reference layout (resnext.py:88-111).

    ``calibrated=False`` is the constructor's raw init (kaiming-normal fan_out convs, unit batch-norm, zero biases),
    which maps every log-mel image to the same class.  ``calibrated=True`` (default; seed 4321 only) jitters the
    batch-norm affine parameters and takes the data-dependent tensors -- batch-norm running statistics measured on
    synthetic log-mels, and a linear layer centred and scaled on the resulting features -- from
    ``data/resnext_calib.npz``, so that predictions on ``clips()`` span the classes, with some near-ties."""
    if calibrated:
        if not (seed == 4321 and nlabels == 10 and in_channels == 1 and depth == 29):
            raise ValueError("data/resnext_calib.npz calibrates the seed-4321 ResNeXt-29 with 10 labels only")
        sd = resnext_state_dict(seed, nlabels, in_channels, cardinality, depth, base_width, widen, calibrated=False)
        rng2 = np.random.Generator(np.random.PCG64(seed + 1))
        for k in [k for k in sd if k.endswith("running_mean")]:
            p = k[:-len(".running_mean")]
            c = sd[k].numel()
            sd[p + ".weight"] = torch.from_numpy(rng2.uniform(0.7, 1.3, size=c).astype(np.float32))
            sd[p + ".bias"] = torch.from_numpy(rng2.normal(0.0, 0.1, size=c).astype(np.float32))
        with np.load(CALIB_FILE) as cal:
            for k in cal.files:
                sd[k] = torch.from_numpy(cal[k].copy())
        return sd
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def conv(name, cout, cin, k):
        std = np.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = torch.from_numpy(rng.normal(0, std, size=(cout, cin, k, k)).astype(np.float32))

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    stages = [64, 64 * widen, 128 * widen, 256 * widen]
    conv("conv_1_3x3", 64, in_channels, 3)
    bn("bn_1", 64)
    for s in (1, 2, 3):
        cin, cout = stages[s - 1], stages[s]
        for j in range((depth - 2) // 9):
            name = "stage_%d.stage_%d_bottleneck_%d" % (s, s, j)
            ci = cin if j == 0 else cout
            D = cardinality * int(base_width * (cout / (widen * 64.0)))
            conv(name + ".conv_reduce", D, ci, 1)
            bn(name + ".bn_reduce", D)
            conv(name + ".conv_conv", D, D // cardinality, 3)
            bn(name + ".bn", D)
            conv(name + ".conv_expand", cout, D, 1)
            bn(name + ".bn_expand", cout)
            if ci != cout:
                conv(name + ".shortcut.shortcut_conv", cout, ci, 1)
                bn(name + ".shortcut.shortcut_bn", cout)
    sd["classifier.weight"] = torch.from_numpy(rng.normal(0, np.sqrt(2.0 / stages[3]), size=(nlabels, stages[3])).astype(np.float32))
    sd["classifier.bias"] = torch.zeros(nlabels)
    return sd


def waveforms(batch, length=16000, seed=0):
    """SURVEY.md section 8d synthetic clips: 0.5*(2U-1), shape (B,1,L) fp32 in [-0.5, 0.5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((0.5 * (2.0 * rng.random((batch, 1, length)) - 1.0)).astype(np.float32))


def clips(batch, length=16000, seed=0, sample_rate=16000):
    """Structured synthetic clips (B,1,L) fp32 in [-0.9, 0.9]: 1-4 gated, slightly chirped partials at log-uniform
    frequencies in [80, 7000] Hz over a noise floor of -60..-20 dB, random peak level.  ``waveforms()`` is white
    noise -- every clip has the same log-mel image up to estimation noise; these differ, so a classifier separates
    them."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(length, dtype=np.float64) / sample_rate
    out = np.zeros((batch, 1, length), dtype=np.float64)
    for b in range(batch):
        y = np.zeros(length, dtype=np.float64)
        for _ in range(int(rng.integers(1, 5))):
            f0 = 80.0 * (7000.0 / 80.0) ** rng.random()
            chirp = rng.uniform(-0.5, 0.5) * f0
            amp = rng.uniform(0.2, 1.0)
            on = rng.uniform(0.0, 0.6)
            dur = rng.uniform(0.15, 0.8)
            env = np.clip((t - on) / 0.02, 0.0, 1.0) * np.clip((on + dur - t) / 0.02, 0.0, 1.0)
            y += amp * env * np.sin(2.0 * np.pi * (f0 * t + 0.5 * chirp * t * t) + rng.uniform(0.0, 2.0 * np.pi))
        y += 10.0 ** rng.uniform(-3.0, -1.0) * rng.standard_normal(length)
        out[b, 0] = y * (rng.uniform(0.2, 0.9) / np.abs(y).max())
    return torch.from_numpy(out.astype(np.float32))


def noise(shape, seed=7):
    """Pre-drawn standard-normal noise (for injected-noise comparisons)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
