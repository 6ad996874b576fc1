"""Parity of the CUDA path (through the C ABI) against the oracle / golden fixtures.  Needs a B200.

Tolerances: tests/gates.py (north_star's waveform rel-L2 <= 1e-2 for bf16 is an upper bound; the gates that bite are
eps with the per-clip mean removed and the waveform error relative to the network's contribution);
log-mel max-abs <= 2e-2 dB against torchaudio; vote counts bit-exact given the same noise; top-1 agreement >= 99.5 %
against the reference's own logits on 256 structured clips.
"""

import ctypes
import json

import numpy as np
import pytest
import torch

import audiopure_b200 as ap
from audiopure_b200 import _lib
from oracle import certify as o_certify, purify as o_purify, resnext as o_resnext, schedule as o_schedule, \
    wavenet as o_wavenet, weights as W
from tests.emulate import emulate_eps
from tests.gates import check_eps, check_wave, margins, rel_l2, zero_eps  # noqa: F401
from tests import gates

pytestmark = pytest.mark.gpu

EPS_GATE = gates.EPS_GATE["bf16"]
WAVE_GATE = gates.WAVE_GATE["bf16"]
SMALL = dict(W.DEFAULT_WAVENET_CONFIG, num_res_layers=6, dilation_cycle=3)


@pytest.fixture(autouse=True, scope="module")
def _true_fp32_consumer():
    """The fp32 ``CifarResNeXt`` module is the consumer the reference runs; keep cuDNN / cuBLAS from silently using
    TF32 for it, so logits differ from the CPU reference by fp32 rounding only (near-tie draws stay decidable)."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def make_model(cfg, seed):
    m = ap.WaveNet_Speech_Commands(**cfg)
    m.load_state_dict(W.make_state_dict(seed, cfg))
    return m.cuda().eval()


@pytest.fixture(scope="module")
def full_model():
    return make_model(W.DEFAULT_WAVENET_CONFIG, 1234)


@pytest.fixture(scope="module")
def small_model():
    return make_model(SMALL, 99)


@pytest.fixture(scope="module")
def hp():
    return ap.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)


@pytest.fixture(scope="module")
def o_hp():
    return o_schedule.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)


@pytest.fixture(scope="module")
def classifier():
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(o_resnext.make_state_dict(4321))
    return clf.cuda().eval()


# ----------------------------------------------------------------------------- bring-up: tcgen05 / TMA --
@pytest.mark.parametrize("K", [64, 256, 768])
def test_debug_gemm(K):
    lib = _lib.load()
    g = torch.Generator().manual_seed(K)
    a = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(256, K, generator=g).to(torch.bfloat16).cuda()
    d = torch.zeros(128, 256, device="cuda")
    _lib.check(lib.ap_debug_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), K, _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = a.float().cpu() @ b.float().cpu().t()
    assert rel_l2(d, want) < 1e-5


# ------------------------------------------------------------------------------------ epsilon network --
def test_eps_small_ragged_vs_emulation_oracle_golden(small_model, golden):
    g = golden("wavenet_small.npz")
    t = int(g["t"])
    x = W.make_waveforms(3, 1000, seed=int(g["x_seed"]))
    got = small_model((x.cuda(), t * torch.ones(3, 1))).cpu()
    assert got.shape == (3, 1, 1000)
    packed = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in small_model.engine().packed.items()}
    emu = emulate_eps(packed, x, t, 6, 3, quantize=True)
    assert rel_l2(got, emu) < 5e-3          # same rounding points: only accumulation order / tanh.approx differ
    check_eps(got, g["eps"])                # reference output


def test_layer_intermediates_vs_emulation(small_model):
    """Every layer's gate tile (read back from the workspace) against the emulation: localises a bad layer,
    including the clip edges and the ragged last tile (L = 1000 = 7*128 + 104)."""
    x = W.make_waveforms(2, 1000, seed=11)
    eng = small_model.engine()
    eng.eps(x.cuda(), 3)
    torch.cuda.synchronize()
    B, L, layers = 2, 1000, 6
    h_bytes = B * L * 256 * 2
    ws = eng.workspace(B, L)
    gate = ws[2 * h_bytes: 2 * h_bytes + layers * h_bytes].view(torch.bfloat16).view(layers, B, L, 256).float().cpu()
    packed = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in eng.packed.items()}
    _, inter = emulate_eps(packed, x, 3, 6, 3, quantize=True, return_inter=True)
    for n in range(layers):
        err = rel_l2(gate[n], inter["gate"][n])
        assert err < 1e-2, "layer %d gate rel-L2 %.3e" % (n, err)
        edge = torch.cat([gate[n][:, :8], gate[n][:, -8:]], 1)
        want = torch.cat([inter["gate"][n][:, :8], inter["gate"][n][:, -8:]], 1)
        assert rel_l2(edge, want) < 2e-2, "layer %d clip edges" % n


@pytest.mark.parametrize("t", [1, 33])
def test_eps_full_vs_golden(full_model, golden, t):
    g = golden("wavenet_full.npz")
    x = W.make_waveforms(1, 16000, seed=0)
    got = full_model.engine().eps(x.cuda(), t)
    tot, acv = check_eps(got, g["eps_t%d" % t])
    print("eps t=%d: rel-L2 %.3e, mean-removed %.3e" % (t, tot, acv))


def test_eps_dilation_2048_layer_vs_golden(full_model, golden):
    """Layers 0, 11 (dilation 2048) and 35 of the full network at the sampled time steps (clip edges included)."""
    g = golden("wavenet_full.npz")
    x = W.make_waveforms(1, 16000, seed=0)
    eng = full_model.engine()
    eng.eps(x.cuda(), 1)
    torch.cuda.synchronize()
    h_bytes = 16000 * 256 * 2
    ws = eng.workspace(1, 16000)
    idx = torch.from_numpy(g["slice_t"]).long()
    sd = W.make_state_dict(1234)
    for n in (0, 11, 35):
        gate = ws[2 * h_bytes + n * h_bytes: 2 * h_bytes + (n + 1) * h_bytes].view(torch.bfloat16).view(16000, 256)
        gate = 0.5 * gate.float().cpu()[idx]                                   # (64, 256); the kernels keep 2 x gate
        ws_n = o_wavenet.fold_weight_norm(sd["residual_layer.residual_blocks.%d.skip_conv.weight_g" % n],
                                          sd["residual_layer.residual_blocks.%d.skip_conv.weight_v" % n])[:, :, 0]
        skip = gate @ ws_n.t() + sd["residual_layer.residual_blocks.%d.skip_conv.bias" % n]
        want = torch.from_numpy(g["skip_%d" % n])[0].t()                        # (64, 256)
        assert rel_l2(skip, want) < EPS_GATE, n


def test_eps_is_batch_invariant_at_full_batch(full_model):
    """BASELINE config 2 size (B = 64): every clip of the batch equals its own B = 1 evaluation bit for bit
    (tiles never mix clips), so the small-case parity carries to the full size."""
    x = W.make_waveforms(64, 16000, seed=2).cuda()
    eng = full_model.engine()
    big = eng.eps(x, 1)
    for i in (0, 31, 63):
        one = eng.eps(x[i:i + 1], 1)
        assert torch.equal(big[i:i + 1], one), i
    assert torch.isfinite(big).all()


def test_eps_odd_length_and_tiny_clip(small_model):
    """Ragged shapes: odd L (tile of 1 valid row at the end), L below one tile, single-clip pair with a dummy tile."""
    sd = W.make_state_dict(99, SMALL)
    for B, L in ((1, 129), (3, 77), (1, 128)):
        x = W.make_waveforms(B, L, seed=L)
        got = small_model.engine().eps(x.cuda(), 4)
        check_eps(got, o_wavenet.eps_theta(sd, x, 4, SMALL), what="eps B=%d L=%d" % (B, L))


@pytest.mark.parametrize("layers,cycle,B,L,t", [
    (1, 1, 1, 64, 0), (2, 2, 2, 127, 5), (3, 3, 5, 300, 17), (5, 5, 1, 2049, 100), (12, 12, 2, 700, 3),
    (12, 12, 1, 4100, 199), (4, 2, 4, 1536, 8), (7, 7, 3, 257, 60),
])
def test_eps_random_architectures_and_shapes(layers, cycle, B, L, t):
    """Depths, dilation cycles (up to dilation 2048 > L: every outer tap window is dead), odd batch sizes and ragged
    lengths against the dataflow emulation -- exercises dummy tiles of odd pairs, partial tiles and skipped taps."""
    cfg = dict(W.DEFAULT_WAVENET_CONFIG, num_res_layers=layers, dilation_cycle=cycle)
    m = make_model(cfg, 1000 + layers)
    x = W.make_waveforms(B, L, seed=L)
    got = m.engine().eps(x.cuda(), t).cpu()
    packed = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in m.engine().packed.items()}
    emu = emulate_eps(packed, x, t, layers, cycle, quantize=True)
    assert got.shape == (B, 1, L) and torch.isfinite(got).all()
    assert rel_l2(got, emu) < 8e-3, rel_l2(got, emu)


def test_eps_chunking_over_max_chunk():
    m = ap.WaveNet_Speech_Commands(**SMALL, max_chunk=2)
    m.load_state_dict(W.make_state_dict(99, SMALL))
    m = m.cuda()
    ref = make_model(SMALL, 99)
    x = W.make_waveforms(5, 512, seed=4).cuda()
    assert torch.equal(m.engine().eps(x, 2), ref.engine().eps(x, 2))


# -------------------------------------------------------------------------------------------- purifiers --
@pytest.mark.parametrize("t_star", [2, 3])
def test_ddpm_purify_vs_reference(full_model, hp, o_hp, golden, t_star):
    g = golden("ddpm_t%d.npz" % t_star)
    x = W.make_waveforms(2, 16000, seed=int(g["x_seed"]))
    z = W.make_noise((t_star, 2, 1, 16000), seed=int(g["z_seed"]))
    dw = ap.DiffWave(full_model, hp, reverse_timestep=t_star)
    y = dw(x.cuda(), z=z)
    assert y.shape == x.shape and y.is_cuda
    tot, net = check_wave(y, g["purified"], o_purify.ddpm_purify(o_hp, zero_eps, x, t_star, z))
    print("ddpm t*=%d: waveform rel-L2 %.3e, error / network contribution %.3e" % (t_star, tot, net))
    # step-by-step surface (compute_coefficients / _diffusion / _reverse) agrees with the fused loop
    y2 = dw._reverse(dw._diffusion(x.cuda(), z=z[0]), z=z[1:])
    assert rel_l2(y2, y) < 1e-5


def test_ddpm_accepts_numpy_and_leaves_input_intact(small_model, hp):
    dw = ap.DiffWave(small_model, hp, reverse_timestep=2)
    x = W.make_waveforms(2, 640, seed=1)
    keep = x.clone()
    xc = x.cuda()
    y = dw(x.numpy())
    assert y.shape == x.shape
    dw(xc)
    assert torch.equal(xc.cpu(), keep)


def test_one_shot_vs_reference(full_model, hp, o_hp, golden):
    g = golden("oneshot_t34.npz")
    x = W.make_waveforms(1, 16000, seed=0)
    dw = ap.DiffWave(full_model, hp, reverse_timestep=int(g["reverse_timestep"]))
    check_wave(dw.one_shot_denoise(x.cuda()), g["x0_hat"], o_purify.one_shot_denoise(o_hp, zero_eps, x, 34))


def test_one_shot_at_the_certifier_sigmas(full_model, hp, o_hp, golden):
    """certified_robust.py:102-110: sigma = 0.1 / 0.5 / 1.0 -> t* = 14 / 66 / 117, on the certifier's kind of input
    sqrt(abar*) (x + sigma z); fixture from the reference's own one_shot_denoise."""
    g = golden("oneshot_sigmas.npz")
    x = W.make_clips(1, 16000, seed=int(g["x_seed"]))
    z = W.make_noise((1, 1, 16000), seed=int(g["z_seed"]))
    for sigma, t_star in zip(g["sigmas"].tolist(), g["t_stars"].tolist()):
        x_t = (1 / (1 + sigma ** 2)) ** 0.5 * (x + sigma * z)
        dw = ap.DiffWave(full_model, hp, reverse_timestep=int(t_star))
        tot, net = check_wave(dw.one_shot_denoise(x_t.cuda()), g["x0_hat_t%d" % t_star],
                              o_purify.one_shot_denoise(o_hp, zero_eps, x_t, int(t_star)), what="one-shot t*=%d" % t_star)
        print("one-shot t*=%d: rel-L2 %.3e, error / network contribution %.3e" % (t_star, tot, net))


def test_compute_coefficients_and_eps_t(full_model, hp, golden):
    g = golden("wavenet_full.npz")
    x = W.make_waveforms(1, 16000, seed=0).cuda()
    dw = ap.DiffWave(full_model, hp, reverse_timestep=2)
    eps, mu, sigma = dw.compute_coefficients(x, 1)
    check_eps(eps, g["eps_t1"])
    a, ab = float(hp["Alpha"][1]), float(hp["Alpha_bar"][1])
    want = (x.cpu() - (1 - a) / (1 - ab) ** 0.5 * torch.from_numpy(g["eps_t1"])) / a ** 0.5
    assert rel_l2(mu, want) < 1e-3
    assert float(sigma) == float(hp["Sigma"][1])
    assert torch.equal(dw.compute_eps_t(x, torch.tensor(1)), eps)


def test_sde_purify_vs_oracle(full_model, hp):
    """torchsde is absent: the Euler-Maruyama restatement (oracle.purify.sde_purify, f/g pinned to the
    reference's RevVPSDE) is the checker."""
    sd = W.make_state_dict(1234)
    tab = o_schedule.sde_tables()
    t = 2
    x = W.make_waveforms(1, 16000, seed=3)
    z = W.make_noise((t + 1, 1, 1, 16000), seed=22)
    want = o_purify.sde_purify(tab, lambda xx, k: o_wavenet.eps_theta(sd, xx, k), x, t, z[0], z[1:].reshape(t, 1, 16000))

    class Args:
        pass

    args = Args()
    args.t, args.sample_step, args.rand_t, args.t_delta, args.use_bm, args.score_type = t, 1, False, 0, False, "guided_diffusion"
    rev = ap.RevDiffWave(args, model=ap.DiffWave(full_model, hp, reverse_timestep=t))
    got = rev(x.cuda(), z=z[None])
    assert got.shape == (1, 1, 16000)
    check_wave(got, want, o_purify.sde_purify(tab, zero_eps, x, t, z[0], z[1:].reshape(t, 1, 16000)))
    # RevVPSDE.f / .g surface against the reference fixture
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "sde_fg.npz"))
    xf = W.make_waveforms(1, 16000, seed=int(g["x_seed"])).view(1, -1).cuda()
    for i, tc in enumerate(g["tc"]):
        tt = torch.tensor(tc, dtype=torch.float32)
        assert rel_l2(rev.rev_vpsde.f(tt, xf), g["f"][i]) < 2e-2
        np.testing.assert_allclose(rev.rev_vpsde.g(tt, xf)[:, :4].cpu().numpy(), g["g"][i], rtol=1e-5, atol=0)


@pytest.mark.parametrize("t,B", [(5, 2), (10, 1)])
def test_sde_purify_more_steps_vs_reference_drift(full_model, hp, golden, t, B):
    """BASELINE configs[2] ("more reverse steps"): full network, t = 5 (B = 2) and t = 10, against fixtures made by
    Euler-Maruyama steps of the reference's OWN RevVPSDE.f / .g (diffwave_sde.py:118-134) on structured clips."""
    g = golden("sde_t%d.npz" % t)
    x = W.make_clips(B, 16000, seed=int(g["x_seed"]))
    z = W.make_noise((t + 1, B, 1, 16000), seed=int(g["z_seed"]))
    args = type("A", (), dict(t=t, sample_step=1, rand_t=False, t_delta=0, use_bm=False, score_type="guided_diffusion"))()
    rev = ap.RevDiffWave(args, model=ap.DiffWave(full_model, hp, reverse_timestep=t))
    got = rev(x.cuda(), z=z[None])
    y0 = o_purify.sde_purify(o_schedule.sde_tables(), zero_eps, x, t, z[0], z[1:].reshape(t, B, 16000))
    tot, net = check_wave(got, g["purified"], y0, what="sde t=%d" % t)
    print("sde t=%d B=%d: waveform rel-L2 %.3e, error / network contribution %.3e" % (t, B, tot, net))


def test_purify_of_130_clips_equals_its_chunks(full_model):
    """B = 130 at L = 16000 runs as chunks of 64 + 64 + 2 (the path configs[2] and [4] take): bit-equal to purifying
    each chunk on its own with the matching clip offset (Philox noise is keyed on the global clip index)."""
    eng = full_model.engine()
    x = W.make_clips(130, 16000, seed=90).cuda()
    whole = eng.ddpm_purify(x, 2, seed=5)
    for lo, hi in ((0, 64), (64, 128), (128, 130)):
        part = eng.ddpm_purify(x[lo:hi], 2, seed=5, clip_offset=lo)
        assert torch.equal(whole[lo:hi], part), (lo, hi)
    sde = eng.sde_purify(x, 1, seed=6)
    assert torch.equal(sde[128:], eng.sde_purify(x[128:], 1, seed=6, clip_offset=128))
    assert torch.isfinite(whole).all() and torch.isfinite(sde).all()


def test_sde_sample_step_concatenates(small_model, hp):
    class Args:
        t, sample_step, rand_t, t_delta, use_bm, score_type = 2, 3, False, 0, False, "guided_diffusion"

    rev = ap.RevDiffWave(Args(), model=ap.DiffWave(small_model, hp, reverse_timestep=2))
    out = rev(W.make_waveforms(2, 512, seed=1).cuda())
    assert out.shape == (6, 1, 512)  # diffwave_sde.py:212


def test_sde_linearised_gradient_matches_oracle_autograd(small_model, hp):
    """with eps under no_grad the purifier's input-Jacobian is a scalar; check it against autograd
    through the oracle's Euler-Maruyama restatement (eps detached, as compute_eps_t does)."""
    sd = W.make_state_dict(99, SMALL)
    tab = o_schedule.sde_tables()
    t = 3
    x = W.make_waveforms(1, 2048, seed=6)
    z = W.make_noise((t + 1, 1, 1, 2048), seed=23)
    xo = x.clone().requires_grad_(True)
    eps_fn = lambda xx, k: o_wavenet.eps_theta(sd, xx.detach(), k, SMALL)  # noqa: E731
    yo = o_purify.sde_purify(tab, eps_fn, xo, t, z[0], z[1:].reshape(t, 1, 2048))
    w = W.make_noise((1, 1, 2048), seed=24)
    (yo * w).sum().backward()

    class Args:
        pass

    args = Args()
    args.t, args.sample_step, args.rand_t, args.t_delta, args.use_bm, args.score_type = t, 1, False, 0, False, "guided_diffusion"
    rev = ap.RevDiffWave(args, model=ap.DiffWave(small_model, hp, reverse_timestep=t))
    xg = x.cuda().requires_grad_(True)
    yg = rev(xg, z=z[None])
    (yg * w.cuda()).sum().backward()
    assert rel_l2(yg, yo) < WAVE_GATE
    assert rel_l2(xg.grad, xo.grad) < 1e-4
    assert abs(rev.input_jacobian(t) - float((xo.grad / w).mean())) < 1e-4


def test_purify_is_cuda_graph_capturable(small_model):
    """The C ABI enqueues on the caller's stream with no host synchronisation and no hidden allocation, so a whole
    purification (diffuse + t* x (prologue, layers, tail), programmatic dependent launches included) captures into
    a CUDA graph; the replay is bit-equal to the eager call."""
    eng = small_model.engine()
    x = W.make_waveforms(2, 2048, seed=4).cuda()
    z = W.make_noise((3, 2, 1, 2048), seed=5).cuda()
    eager = eng.ddpm_purify(x, 3, z=z)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.ddpm_purify(x, 3, z=z)  # workspace + tensor maps exist before capture
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        out = eng.ddpm_purify(x, 3, z=z)
    x.copy_(W.make_waveforms(2, 2048, seed=6))  # new input in the captured buffer
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eng.ddpm_purify(x, 3, z=z))
    assert not torch.equal(out, eager)


# ------------------------------------------------------------------------------------------ philox noise --
def test_philox_noise_is_shard_invariant_and_seeded(small_model, hp):
    dw = ap.DiffWave(small_model, hp, reverse_timestep=3, seed=5)
    x = W.make_waveforms(4, 1024, seed=8).cuda()
    eng = small_model.engine()
    whole = eng.ddpm_purify(x, 3, seed=123)
    again = eng.ddpm_purify(x, 3, seed=123)
    assert torch.equal(whole, again)
    halves = torch.cat([eng.ddpm_purify(x[:2], 3, seed=123, clip_offset=0),
                        eng.ddpm_purify(x[2:], 3, seed=123, clip_offset=2)])
    assert torch.equal(whole, halves)            # sharding the batch does not change the draws
    other = eng.ddpm_purify(x, 3, seed=124)
    assert not torch.equal(whole, other)
    assert not torch.equal(dw(x), dw(x))         # successive forward() calls draw fresh noise


def test_philox_normal_moments():
    lib = _lib.load()
    L, n = 4096, 256
    x = torch.zeros(1, L, device="cuda")
    out = torch.empty(n, 1, L, device="cuda")
    _lib.check(lib.ap_smooth_inputs(x.data_ptr(), L, n, 1.0, 1.0, None, 42, 3, 0, out.data_ptr(), _lib.stream_ptr()))
    v = out.double().flatten()
    assert abs(float(v.mean())) < 5e-3 and abs(float(v.var()) - 1) < 1e-2
    assert abs(float((v ** 4).mean()) - 3) < 0.05
    # draws are keyed on the draw index: asking for draws [100, 110) reproduces that slice
    sub = torch.empty(10, 1, L, device="cuda")
    _lib.check(lib.ap_smooth_inputs(x.data_ptr(), L, 10, 1.0, 1.0, None, 42, 3, 100, sub.data_ptr(), _lib.stream_ptr()))
    assert torch.equal(sub, out[100:110])


# ------------------------------------------------------------------------------------------- front-end --
def test_logmel_vs_torchaudio_fixture(golden):
    g = golden("mel.npz")
    x = torch.cat([W.make_waveforms(2, 16000, seed=0), torch.from_numpy(golden("ddpm_t2.npz")["purified"]),
                   W.make_clips(4, 16000, seed=20)], 0)
    tr = ap.LogMelSpectrogram().cuda()
    got = tr(x.cuda())
    assert got.shape == (8, 1, 32, 32)
    assert float((got.cpu() - torch.from_numpy(g["logmel"])).abs().max()) < 2e-2


def test_logmel_ragged_lengths_and_silence():
    from oracle import mel as o_mel

    tr = ap.LogMelSpectrogram().cuda()
    for L in (512, 8000, 16384 + 300):
        x = W.make_waveforms(3, L, seed=L)
        want = o_mel.log_mel(x)
        got = tr(x.cuda())
        assert got.shape == want.shape
        assert float((got.cpu() - want).abs().max()) < 2e-2
    z = tr(torch.zeros(1, 1, 16000, device="cuda"))
    assert float(z.max()) == -100.0 and float(z.min()) == -100.0  # clamp at 1e-10 -> -100 dB


def test_logmel_large_batch_equals_per_clip_runs():
    """130 clips = 2080 frame pairs > the 1036 resident CTAs: the persistent loop of the log-mel kernel re-uses its
    shared-memory buffers for several items; every clip must equal its own single-clip launch bit for bit, and the
    oracle within the dB gate."""
    from oracle import mel as o_mel

    tr = ap.LogMelSpectrogram().cuda()
    x = W.make_clips(130, 16000, seed=77).cuda()
    big = tr(x)
    for i in (0, 1, 64, 65, 129):
        assert torch.equal(big[i:i + 1], tr(x[i:i + 1])), i
    assert float((big[100:104].cpu() - o_mel.log_mel(x[100:104].cpu())).abs().max()) < 2e-2


def test_logmel_backward_vs_autograd_of_oracle():
    from oracle import mel as o_mel

    tr = ap.LogMelSpectrogram().cuda()
    for L in (16000, 5000):
        x = W.make_waveforms(2, L, seed=L + 1)
        w = W.make_noise((2, 1, 32, 1 + L // 512), seed=3)
        xo = x.clone().requires_grad_(True)
        (o_mel.log_mel(xo) * w).sum().backward()
        xg = x.cuda().requires_grad_(True)
        (tr(xg) * w.cuda()).sum().backward()
        assert rel_l2(xg.grad, xo.grad) < 1e-3, L


def test_gradient_flows_through_acoustic_system(small_model, hp, classifier):
    """The adaptive attack's backward pass (white_box_attack.py:392,437-439): loss -> classifier -> log-mel kernel
    -> SDE purifier (linearised Jacobian) -> perturbation."""

    class Args:
        t, sample_step, rand_t, t_delta, use_bm, score_type = 2, 1, False, 0, False, "guided_diffusion"

    rev = ap.RevDiffWave(Args(), model=ap.DiffWave(small_model, hp, reverse_timestep=2))
    AS = ap.AcousticSystem(classifier=classifier, transform=ap.LogMelSpectrogram().cuda(), defender=rev)
    x = W.make_waveforms(2, 16000, seed=13).cuda()
    delta = torch.zeros_like(x, requires_grad=True)
    loss = torch.nn.functional.cross_entropy(AS(x + delta), torch.tensor([1, 2], device="cuda"))
    loss.backward()
    assert delta.grad is not None and torch.isfinite(delta.grad).all() and float(delta.grad.abs().max()) > 0


# ------------------------------------------------------------------------- composition and certification --
def test_acoustic_system_vs_reference(full_model, hp, classifier, golden):
    g = golden("acoustic.npz")
    x = W.make_clips(2, 16000, seed=int(g["x_seed"])).cuda()
    z = W.make_noise((2, 2, 1, 16000), seed=int(g["z_seed"]))
    dw = ap.DiffWave(full_model, hp, reverse_timestep=2)

    class Injected(torch.nn.Module):  # the defender slot is any callable (B,1,L)->(B,1,L)
        def forward(self, w):
            return dw(w, z=z)

    AS = ap.AcousticSystem(classifier=classifier, transform=ap.LogMelSpectrogram().cuda(), defender=Injected())
    with torch.no_grad():
        logits = AS(x)
        nodef = AS(x, defend=False)
    assert rel_l2(nodef, g["logits_nodefend"]) < 1e-3
    assert rel_l2(logits, g["logits"]) < 1e-2
    assert np.array_equal(logits.argmax(1).cpu().numpy(), g["logits"].argmax(1))
    assert not np.array_equal(g["logits"].argmax(1), g["logits_nodefend"].argmax(1))  # the defence changes the answer
    with pytest.raises(NotImplementedError):
        ap.AcousticSystem(classifier, None, dw, defense_type="other")


def test_pipeline_top1_agreement_on_256_clips_vs_reference(full_model, hp, classifier, golden):
    """north_star: classifier top-1 agreement >= 99.5 % on synthetic clips.  BASELINE configs[1] workload (DDPM t*=2 ->
    log-mel -> ResNeXt-29) on 256 structured clips with injected noise, against the logits and the sampled purified
    waveforms the UNMODIFIED reference produced for them (tests/golden/pipeline256.npz): the calibrated classifier
    spreads these clips over >= 5 classes, some with margins below 0.05, so a purifier or front-end bug shows."""
    g = golden("pipeline256.npz")
    want = torch.from_numpy(g["logits"])
    want_pred = want.argmax(1)
    assert len(set(want_pred.tolist())) >= 5
    B, t_star = want.shape[0], int(g["t_star"])
    x = W.make_clips(B, 16000, seed=int(g["x_seed"]))
    z = W.make_noise((t_star, B, 1, 16000), seed=int(g["z_seed"]))
    dw = ap.DiffWave(full_model, hp, reverse_timestep=t_star)
    AS = ap.AcousticSystem(classifier, ap.LogMelSpectrogram().cuda(), defender=None)
    with torch.no_grad():
        wave = dw(x.cuda(), z=z)
        logits = AS(wave).cpu()
    idx = torch.from_numpy(g["sample_idx"]).long()
    o_hp = o_schedule.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)
    y0 = o_purify.ddpm_purify(o_hp, zero_eps, x, t_star, z)[:, 0, idx]
    tot, net = check_wave(wave[:, 0, idx.cuda()], g["purified_samples"], y0)
    agree = (logits.argmax(1) == want_pred)
    m = margins(want)
    near = m < 0.05
    print("pipeline256: waveform rel-L2 %.3e (error / network contribution %.3e); logits rel-L2 %.3e; top-1 agreement "
          "%.4f; %d near-ties (margin < 0.05, smallest %.2e) of which %d agree; classes %s"
          % (tot, net, rel_l2(logits, want), float(agree.float().mean()), int(near.sum()), float(m.min()),
             int(agree[near].sum()), sorted(set(want_pred.tolist()))))
    assert int(near.sum()) >= 2
    assert rel_l2(logits, want) < 1e-2
    assert float(agree.float().mean()) >= 0.995
    assert bool(agree[m > 0.02].all())        # away from near-ties every prediction matches


@pytest.mark.parametrize("with_res,relu", [(False, True), (True, True), (True, False)])
def test_bias_act_epilogue_kernel_bit_exact(with_res, relu):
    """ap_bias_act_nhwc_bf16: y <- relu?(y + bias[c] (+ res)) in place over channels-last bf16, fp32 arithmetic, one
    rounding -- bit-exact against the same expression in torch (ragged row count, C not a power of two)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    n, c, h, w = 3, 72, 5, 7
    y = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    res = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(c, generator=g).cuda()
    want = y.float() + bias.reshape(1, -1, 1, 1)
    if with_res:
        want = want + res.float()
    if relu:
        want = want.relu()
    want = want.to(torch.bfloat16)
    _lib.check(lib.ap_bias_act_nhwc_bf16(y.data_ptr(), bias.data_ptr(), res.data_ptr() if with_res else None,
                                         n * h * w, c, int(relu), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(y, want)
    with pytest.raises(_lib.AudioPureError):
        _lib.check(lib.ap_bias_act_nhwc_bf16(y.data_ptr(), bias.data_ptr(), None, 10, 12, 1, _lib.stream_ptr()))


def test_fused_bf16_classifier_agrees_with_fp32(classifier):
    """The bf16 inference form of the consumer against the fp32 module on 512 structured clips.  bf16
    logits carry ~1e-2 relative error, so near-ties (fp32 margin below that) can legitimately flip: agreement must be
    total away from them and is reported on them."""
    fused = ap.FusedResNeXt(classifier).cuda()
    tr = ap.LogMelSpectrogram().cuda()
    x = W.make_clips(512, 16000, seed=12).cuda()
    with torch.no_grad():
        spec = tr(x)
        a, b = classifier(spec), fused(spec)
    m = margins(a.cpu())
    agree = (a.argmax(1) == b.argmax(1)).cpu()
    print("fused bf16 vs fp32: logits rel-L2 %.3e; agreement %.4f overall, %d / %d on margins < 0.25"
          % (rel_l2(b, a), float(agree.float().mean()), int(agree[m < 0.25].sum()), int((m < 0.25).sum())))
    assert len(set(a.argmax(1).tolist())) >= 5
    assert rel_l2(b, a) < 5e-2
    assert bool(agree[m >= 0.25].all())
    assert float(agree.float().mean()) >= 0.95


def test_vote_counts_kernel_bit_exact():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1000, 10, generator=g)
    logits[::7, 3] = logits[::7, 5] = 9.0  # ties: lowest index wins, like torch.max
    counts = torch.zeros(10, dtype=torch.int64, device="cuda")
    lc = logits.cuda()
    _lib.check(lib.ap_vote_counts(lc.data_ptr(), 1000, 10, counts.data_ptr(), _lib.stream_ptr()))
    assert torch.equal(counts.cpu(), o_certify.vote_counts(logits, 10))
    assert int(counts.sum()) == 1000


def _certify_fixture_inputs(g):
    n_0, n = int(g["n_0"]), int(g["n"])
    x = W.make_clips(8, 16000, seed=int(g["x_seed"]))[g["clips"]]
    z = W.make_noise((len(g["clips"]), n_0 + n, 1, 16000), seed=int(g["z_seed"]))
    return x, z, n_0, n


def test_smooth_predict_counts_vs_reference(full_model, hp, classifier, golden):
    """north_star: vote counts bit-exact given the same noise.  128 draws of the estimation pass of clip 0 of
    certify.npz (the unmodified reference's RobustCertificate, injected noise): a multi-class vote histogram."""
    g = golden("certify.npz")
    x, z, n_0, n = _certify_fixture_inputs(g)
    want = g["counts"][0]
    assert int((want > 0).sum()) >= 3 and int(want.sum()) == n == 128
    dw = ap.DiffWave(full_model, hp, reverse_timestep=2)
    RC = ap.RobustCertificate(classifier=classifier, transform=ap.LogMelSpectrogram().cuda(), denoiser=dw)
    counts = RC.smooth_predict(x[0].cuda(), num_sampling=n, sigma=float(g["sigma"]), batch_size=int(g["batch_size"]),
                               z=z[0, n_0:])
    assert dw.reverse_timestep == int(g["t_star"])  # the certifier retargets the denoiser (certified_robust.py:53)
    assert counts.dtype == torch.int64 and not counts.is_cuda
    assert np.array_equal(counts.numpy(), want), (counts.tolist(), want.tolist())


def test_certify_vs_reference_with_abstention(full_model, hp, classifier, golden):
    """certified_robust.py:69-100 on the two fixture clips (n_0 = 32, n = 128, sigma = 0.25, injected noise): per-draw
    logits, both passes' vote counts (bit-exact), and (y_pred, radius) -- one clip certifies, the other ABSTAINS
    (-1, 0) -- against the reference's own outputs and against the oracle's certify_from_counts."""
    g = golden("certify.npz")
    x, z, n_0, n = _certify_fixture_inputs(g)
    sigma, alpha, bs = float(g["sigma"]), float(g["alpha"]), int(g["batch_size"])
    dw = ap.DiffWave(full_model, hp, reverse_timestep=2)
    tr = ap.LogMelSpectrogram().cuda()
    RC = ap.RobustCertificate(classifier=classifier, transform=tr, denoiser=dw)
    y = torch.zeros(2, dtype=torch.long, device="cuda")
    y_pred, radius = RC.certify(x.cuda(), y, sigma=sigma, n_0=n_0, n=n, alpha=alpha, batch_size=bs, z=z)
    c0, c = RC.last_counts
    assert np.array_equal(c0.numpy(), g["counts_0"]), (c0.tolist(), g["counts_0"].tolist())
    assert np.array_equal(c.numpy(), g["counts"]), (c.tolist(), g["counts"].tolist())
    assert np.array_equal(y_pred.cpu().numpy(), g["y_pred"])
    assert sorted(g["y_pred"].tolist())[0] == -1 and sorted(g["y_pred"].tolist())[1] >= 0   # abstain + certify
    np.testing.assert_allclose(radius.cpu().numpy(), g["radius"], rtol=1e-6, atol=0)
    for i in range(2):
        cls, r = o_certify.certify_from_counts(c0[i], c[i], n, sigma, alpha)
        assert (cls, np.float32(r)) == (int(y_pred[i]), np.float32(radius[i].item()))
    # per-draw logits of the work list, batch by batch, against the reference's
    ab = 1 / (1 + sigma ** 2)
    x_in = ab ** 0.5 * (x[:, None].cuda() + sigma * z.cuda()).reshape(-1, 1, 16000)
    with torch.no_grad():
        logits = torch.cat([RC.forward(x_in[s:s + bs]) for s in range(0, x_in.shape[0], bs)]).cpu().reshape(2, n_0 + n, 10)
    want = torch.from_numpy(g["logits"])
    m = margins(want.reshape(-1, 10))
    print("certify: logits rel-L2 %.3e vs the reference; smallest top-2 margin over %d draws %.3e"
          % (rel_l2(logits, want), m.numel(), float(m.min())))
    assert rel_l2(logits, want) < 1e-3


def test_certify_batched_work_list_equals_per_clip_loops(small_model, hp, classifier):
    """The batched certify (one clip-major work list, full batches spanning clips, one read-back) against the
    per-clip form of the reference loop (certified_robust.py:81-93: smooth_predict(n_0) then smooth_predict(n) for
    each clip) with the same Philox keys: bit-equal counts, and the certificate the oracle derives from them."""
    dw = ap.DiffWave(small_model, hp, reverse_timestep=2)
    tr = ap.LogMelSpectrogram().cuda()
    x = W.make_clips(5, 16000, seed=9).cuda()
    y = torch.zeros(5, dtype=torch.long, device="cuda")
    n_0, n, sigma = 10, 50, 0.25
    RC = ap.RobustCertificate(classifier, tr, dw, seed=1)
    y_pred, radius = RC.certify(x, y, sigma=sigma, n_0=n_0, n=n, batch_size=16, clip_offset=100)
    c0, c = RC.last_counts
    assert c0.sum(1).tolist() == [n_0] * 5 and c.sum(1).tolist() == [n] * 5
    loop = ap.RobustCertificate(classifier, tr, dw, seed=1)
    for i in range(5):
        a = loop.smooth_predict(x[i], n_0, sigma, batch_size=7, clip=100 + i, first_draw=0)
        b = loop.smooth_predict(x[i], n, sigma, batch_size=64, clip=100 + i, first_draw=n_0)
        assert torch.equal(a, c0[i]) and torch.equal(b, c[i]), i
        cls, r = o_certify.certify_from_counts(a, b, n, sigma, 0.001)
        assert (cls, np.float32(r)) == (int(y_pred[i]), np.float32(radius[i].item()))
    assert len(set(c.argmax(1).tolist())) >= 2 or int((c > 0).sum()) > 5     # votes are not one class for all
    # successive calls advance the clip keys: fresh, independent draws
    RC2 = ap.RobustCertificate(classifier, tr, dw, seed=1)
    first = RC2.smooth_predict(x[0], 40, sigma)
    second = RC2.smooth_predict(x[0], 40, sigma)
    again = ap.RobustCertificate(classifier, tr, dw, seed=1).smooth_predict(x[0], 40, sigma)
    assert torch.equal(first, again)
    RC2.certify(x[:2], y[:2], sigma=sigma, n_0=4, n=8, batch_size=16)
    assert RC2._next_clip_key == 4


def test_randsmooth_baseline_without_denoiser(classifier):
    """certified_robustness_eval.py:88-89 (`--defense_method randsmooth`): no denoiser, no sqrt(alpha_bar) scaling."""
    tr = ap.LogMelSpectrogram().cuda()
    from oracle import mel as o_mel

    x = W.make_clips(1, 16000, seed=5)[0].cuda()
    z = W.make_noise((48, 1, 16000), seed=6)
    RC = ap.RobustCertificate(classifier, tr, denoiser=None)
    counts = RC.smooth_predict(x, 48, 0.5, batch_size=20, z=z)
    x_in = x.cpu().repeat(48, 1, 1) + 0.5 * z
    want_logits = o_resnext.forward(o_resnext.make_state_dict(4321), o_mel.log_mel(x_in))
    want = o_certify.vote_counts(want_logits, 10)
    print("randsmooth: counts %s, smallest top-2 margin %.3e" % (want.tolist(), float(margins(want_logits).min())))
    assert torch.equal(counts, want)


def test_sharded_counts_equal_single_rank(small_model, hp, classifier):
    """Two logical ranks on one GPU: disjoint draw slices, summed counts == the unsharded counts."""
    dw = ap.DiffWave(small_model, hp, reverse_timestep=2)
    tr = ap.LogMelSpectrogram().cuda()
    x = W.make_clips(1, 16000, seed=4)[0].cuda()
    whole = ap.RobustCertificate(classifier, tr, dw, seed=3).smooth_predict(x, 37, 0.25, batch_size=8)
    parts = []
    for r in range(2):
        rc = ap.RobustCertificate(classifier, tr, dw, seed=3, rank=r, world_size=2, allreduce=lambda c: c)
        parts.append(rc.smooth_predict(x, 37, 0.25, batch_size=8))   # every rank must use the same batch size
    assert int(whole.sum()) == 37
    assert torch.equal(parts[0] + parts[1], whole)
    # the same through certify with 3 logical ranks and several clips: sharding the flattened work list
    xs = W.make_clips(3, 16000, seed=14).cuda()
    y = torch.zeros(3, dtype=torch.long, device="cuda")
    one = ap.RobustCertificate(classifier, tr, dw, seed=3)
    one.certify(xs, y, n_0=5, n=21, batch_size=8)
    acc = [torch.zeros(3, 10, dtype=torch.int64), torch.zeros(3, 10, dtype=torch.int64)]
    for r in range(3):
        rc = ap.RobustCertificate(classifier, tr, dw, seed=3, rank=r, world_size=3, allreduce=lambda c: c)
        rc.certify(xs, y, n_0=5, n=21, batch_size=8)
        acc[0] += rc.last_counts[0]
        acc[1] += rc.last_counts[1]
    assert torch.equal(acc[0], one.last_counts[0]) and torch.equal(acc[1], one.last_counts[1])


def test_factory_reads_reference_config_and_checkpoint(tmp_path):
    """create_diffwave_model (diffwave_ddpm.py:395-411): same config.json keys, same {'model_state_dict': ...} .pkl,
    and a KWS-style non-16000 clip length (kws_adaptive_attack_eval.py:178)."""
    cfgp, ckpt = tmp_path / "config.json", tmp_path / "1000.pkl"
    cfgp.write_text(json.dumps({"diffusion_config": W.DEFAULT_DIFFUSION_CONFIG, "wavenet_config": SMALL,
                                "train_config": {}, "dist_config": {}}))
    sd = W.make_state_dict(99, SMALL)
    torch.save({"model_state_dict": sd, "optimizer_state_dict": {}}, ckpt)
    dw = ap.create_diffwave_model(str(ckpt), str(cfgp), reverse_timestep=2)
    assert dw.reverse_timestep == 2 and dw.diffusion_hyperparams["T"] == 200
    x = W.make_waveforms(2, 12800, seed=3)
    got = dw.compute_eps_t(x.cuda(), 5)
    assert rel_l2(got, o_wavenet.eps_theta(sd, x, 5, SMALL)) < EPS_GATE
    args = type("A", (), dict(ddpm_path=str(ckpt), ddpm_config=str(cfgp), t=2, sample_step=1, rand_t=False, t_delta=0,
                              use_bm=False, score_type="guided_diffusion"))()
    rev = ap.RevDiffWave(args)
    rev.rev_vpsde.audio_shape = (1, 12800)
    assert rev(x.cuda()).shape == (2, 1, 12800)
    with pytest.raises(NotImplementedError):
        ap.RevVPSDE(dw, score_type="score_sde")


# ---------------------------------------------------------------------------------------- error behaviour --
def test_errors_are_loud(small_model, hp):
    lib = _lib.load()
    eng = small_model.engine()
    with pytest.raises(_lib.AudioPureError):
        eng.eps(torch.zeros(1, 1, 512, device="cuda"), 200)          # t out of range
    with pytest.raises(_lib.AudioPureError):
        eng.ddpm_purify(torch.zeros(1, 1, 512, device="cuda"), 0)    # t* out of range
    with pytest.raises(_lib.AudioPureError):
        _lib.check(lib.ap_eps(eng.handle, None, 1, 512, 0, None, None, 0, None))
    with pytest.raises(_lib.AudioPureError):
        ap.LogMelSpectrogram()(torch.zeros(1, 1, 16000))             # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        ap.DiffWave(small_model, hp)(torch.zeros(4, 16000, device="cuda"))  # ndim == 3 (diffwave_ddpm.py:62)
