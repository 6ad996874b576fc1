"""Oracle tooling (test infrastructure): deterministic random-init checkpoints.

No trained DiffWave checkpoint is reachable offline, so tests, fixtures and
the benchmark all use random-init weights of the named architecture.  This
module builds them in the reference's *state-dict layout*
(``WaveNet_Speech_Commands.state_dict()``: 408 tensors for the shipped config;
weight-normed convs stored as ``...conv.weight_g`` (Cout,1,1) +
``...conv.weight_v`` (Cout,Cin,k) + ``bias`` -- WaveNet.py:23-34,57-73,
138-162) from a numpy ``PCG64`` stream, so the same seed gives the same
checkpoint on every machine with no dependency on the reference code or on
torch's RNG.  ``oracle/make_golden.py`` loads these into the real reference
module with ``load_state_dict`` to produce the fixtures.

Magnitudes follow torch's default ``Conv1d``/``Linear`` init (U(+-1/sqrt(fan_in)),
which is what the reference's constructors leave in ``weight_v`` -- the
``kaiming_normal_`` at WaveNet.py:29 lands on the hook-recomputed ``.weight``
and is discarded at the first forward).  ``weight_g`` is ||v|| times a
U(0.9,1.1) factor so that folding weight-norm is actually exercised, and the
zero-initialised output conv (``ZeroConv1d``, WaveNet.py:39-44) is
re-randomised N(0,0.05^2)/N(0,0.01^2) as SURVEY.md section 0 fact 2 requires
(otherwise eps == 0 and parity is vacuous).
"""

from collections import OrderedDict

import numpy as np
import torch

DEFAULT_WAVENET_CONFIG = {
    "in_channels": 1,
    "res_channels": 256,
    "skip_channels": 256,
    "out_channels": 1,
    "num_res_layers": 36,
    "dilation_cycle": 12,
    "diffusion_step_embed_dim_in": 128,
    "diffusion_step_embed_dim_mid": 512,
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


def make_state_dict(seed=1234, wavenet_config=None):
    """Reference-layout state dict (same keys, shapes and order as the reference
    module's own ``state_dict()``; checked in tests/test_oracle_golden.py)."""
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
    sd["final_conv.2.conv.bias"] = torch.from_numpy(
        rng.normal(0.0, 0.01, size=(cfg["out_channels"],)).astype(np.float32))
    return sd


def make_waveforms(batch, length=16000, seed=0):
    """SURVEY.md section 8d synthetic clips: 0.5*(2U-1), shape (B,1,L) fp32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((0.5 * (2.0 * rng.random((batch, 1, length)) - 1.0)).astype(np.float32))


def make_clips(batch, length=16000, seed=0, sample_rate=16000):
    """Structured synthetic clips (B,1,L) fp32 in [-0.9, 0.9]: 1-4 gated, slightly chirped partials at log-uniform
    frequencies in [80, 7000] Hz over a noise floor of -60..-20 dB, random peak level.  Unlike ``make_waveforms``
    (white noise: every clip has the same log-mel image up to estimation noise) these give the classifier
    inputs that differ, so top-1 / vote-count comparisons can fail.  Same generator as
    ``audiopure_b200.synthetic.clips`` (kept in step by tests/test_packing.py)."""
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


def make_noise(shape, seed=7):
    """Pre-drawn standard normal noise to inject into both implementations."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal(shape).astype(np.float32))


def fingerprint(sd):
    """Small checksum of a state dict (stored in the fixtures)."""
    acc = []
    for k, v in sd.items():
        acc.append(float(v.double().sum()))
        acc.append(float(v.double().abs().sum()))
    return np.asarray(acc, dtype=np.float64)
