"""the batched NES / EOT query loops (audiopure_b200/blackbox.py, ap_nes_inputs / ap_nes_grad
through the C ABI) against the reference's own classes (tests/golden/nes.npz, from robustness_eval/_NES.py and
_EOT.py run unmodified) and against the oracle restatement.  Needs a B200."""

import numpy as np
import pytest
import torch

import audiopure_b200 as ap
from audiopure_b200 import _lib
from audiopure_b200.blackbox import EOT, NES
from oracle import blackbox as o_bb, weights as W
from tests.gates import rel_l2

pytestmark = pytest.mark.gpu

LOSS = torch.nn.CrossEntropyLoss(reduction="none")  # _utils.py:116-117 (task 'SCR')


def test_nes_vs_reference_fixture(golden):
    g = golden("nes.npz")
    N, spd, S = int(g["N"]), int(g["samples_per_draw"]), int(g["samples_per_draw_batch"])
    x = W.make_clips(3, N, seed=int(g["x_seed"])).cuda()
    z = W.make_noise((spd // S, 3, S // 2, 1, N), seed=int(g["z_seed"]))
    eot = EOT(o_bb.toy_model(N, seed=int(g["model_seed"])), LOSS, EOT_size=int(g["EOT_size"]),
              EOT_batch_size=int(g["EOT_batch_size"]), use_grad=False)
    nes = NES(spd, S, float(g["sigma"]), eot)
    out = nes(x, torch.from_numpy(g["y"]), z=z)
    for got, key in zip(out[:4], ("mean_loss", "grad", "adver_loss", "adver_score")):
        assert rel_l2(got, g[key]) < 1e-5, key
    assert np.array_equal(out[4], g["predict"])


def _recover_noise(x, S, num_batches, sigma, seed_obj, L):
    """The Philox halves NES's next call will draw, read back through ap_nes_inputs itself (sigma = 1, x = 0)."""
    lib = _lib.load()
    calls = seed_obj._calls + 1
    seed = (seed_obj.seed * 0x9E3779B97F4A7C15 + calls) & 0xFFFFFFFFFFFFFFFF
    zs = []
    zero = torch.zeros_like(x)
    for i in range(num_batches):
        out = torch.empty(x.shape[0] * S, 1, L, device="cuda")
        _lib.check(lib.ap_nes_inputs(zero.data_ptr(), x.shape[0], L, S, 0, 1.0, None, seed, 0, i * (S // 2),
                                     out.data_ptr(), _lib.stream_ptr()))
        out = out.view(x.shape[0], S, 1, L)
        assert torch.equal(out[:, :S // 2], -out[:, S // 2:])     # antithetic pairs (_NES.py:21)
        zs.append(out[:, :S // 2].clone())
    return torch.stack(zs)


def test_nes_philox_noise_vs_oracle_and_sharding():
    """In-kernel Philox draws: recover them, feed them to the oracle restatement of _NES.py, compare; odd N exercises
    the scalar path.  Moments of the draws are checked, successive calls differ, and three logical ranks (draw batches
    dealt out, gradient summed) reproduce the single-rank estimate."""
    N, spd, S, sigma = 1001, 24, 6, 0.02
    x = W.make_clips(2, N, seed=3).cuda()
    y = torch.tensor([2, 5])
    model = o_bb.toy_model(N, seed=5)
    eot = EOT(model, LOSS, EOT_size=1, EOT_batch_size=1, use_grad=False)
    nes = NES(spd, S, sigma, eot, seed=9)
    z = _recover_noise(x, S, spd // S, sigma, nes, N)
    v = z.double().flatten()
    assert abs(float(v.mean())) < 2e-2 and abs(float(v.var()) - 1) < 3e-2
    out = nes(x, y)
    want = o_bb.nes(model, LOSS, x.cpu(), y, z.cpu(), spd, S, sigma)
    for got, w, key in zip(out[:4], want[:4], ("mean_loss", "grad", "adver_loss", "adver_score")):
        assert rel_l2(got, w) < 1e-4, key
    assert np.array_equal(out[4], want[4])
    assert not torch.equal(nes(x, y)[1], out[1])                   # fresh draws every call
    # sharded: same seed and call count on every rank
    single = NES(spd, S, sigma, eot, seed=9)(x, y)
    grad = torch.zeros_like(single[1])
    loss = torch.zeros_like(single[0])
    for r in range(3):
        part = NES(spd, S, sigma, eot, seed=9, rank=r, world_size=3, allreduce=lambda t: t)(x, y)
        grad += part[1]
        loss += part[0]
        assert torch.equal(part[2], single[2]) and torch.equal(part[3], single[3])   # clean query on every rank
    assert rel_l2(grad, single[1]) < 1e-5 and rel_l2(loss, single[0]) < 1e-5


def test_eot_batches_all_repetitions_and_matches_the_loop():
    N = 512
    x = W.make_clips(3, N, seed=8).cuda()
    y = torch.tensor([0, 3, 9]).cuda()
    model = o_bb.toy_model(N, seed=6)
    scores, loss, grad, decisions = EOT(model, LOSS, EOT_size=7, EOT_batch_size=2, use_grad=False)(x, y)
    ws, wl, wd = o_bb.eot(model, LOSS, x.cpu(), y.cpu(), 7, 2)
    assert grad is None and rel_l2(scores, ws) < 1e-5 and rel_l2(loss, wl) < 1e-5
    assert [len(d) for d in decisions] == [6, 6, 6] and [list(map(int, d)) for d in decisions] == [list(map(int, d)) for d in wd]
    # use_grad: d(loss)/dx averaged over the repetitions (_EOT.py:44-55)
    _, _, g2, _ = EOT(model, LOSS, EOT_size=4, EOT_batch_size=2, use_grad=True)(x, y)
    xr = x.clone().requires_grad_(True)
    LOSS(model(xr), y).sum().backward()
    assert rel_l2(g2, xr.grad) < 1e-5


def test_nes_through_the_purification_path():
    """NES over the real AcousticSystem (reduced-depth DiffWave DDPM t*=1 -> log-mel -> ResNeXt): shapes, finiteness,
    and that the clean query's prediction is the system's own."""
    from oracle import resnext as o_resnext

    cfg = dict(W.DEFAULT_WAVENET_CONFIG, num_res_layers=6, dilation_cycle=3)
    m = ap.WaveNet_Speech_Commands(**cfg)
    m.load_state_dict(W.make_state_dict(99, cfg))
    dw = ap.DiffWave(m.cuda().eval(), ap.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG), reverse_timestep=1, seed=4)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(o_resnext.make_state_dict(4321))
    AS = ap.AcousticSystem(clf.cuda().eval(), ap.LogMelSpectrogram().cuda(), defender=dw)
    x = W.make_clips(2, 16000, seed=2).cuda()
    y = torch.tensor([1, 2])
    nes = NES(20, 10, 0.001, EOT(AS, LOSS, EOT_size=2, EOT_batch_size=1, use_grad=False), seed=3)
    mean_loss, grad, adver_loss, adver_score, predict = nes(x, y)
    assert grad.shape == x.shape and torch.isfinite(grad).all() and float(grad.abs().max()) > 0
    assert mean_loss.shape == (2,) and adver_loss.shape == (2,) and adver_score.shape == (2, 10) and predict.shape == (2,)
