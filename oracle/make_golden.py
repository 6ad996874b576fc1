"""Oracle tooling (test infrastructure): generate tests/golden/*.npz from the
UNMODIFIED reference, executed on the CPU in the build container.

    python -m oracle.make_golden            # writes tests/golden/

The reference has no golden vectors of its own, so these
fixtures -- outputs of the reference's own modules on seeded inputs, weights
and injected noise -- are what pins the oracle (tests/test_oracle_golden.py)
and, through it, the CUDA path.  ``/root/reference`` is only read here; the
fixtures travel, the reference does not.

Noise injection: the reference draws with ``torch.normal(mean, std, size=...)``
on the CPU (diffwave_ddpm.py:66,100; certified_robust.py:47).  While a
reference call runs, ``torch.normal`` is replaced by a function that returns
``mean + std * z`` for the next pre-drawn ``z`` (numpy PCG64, oracle.weights.
make_noise), so the unmodified reference code consumes known noise in its own
draw order.
"""

import contextlib
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import _refharness, resnext as o_resnext, weights as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


@contextlib.contextmanager
def injected_normal(z_list):
    queue = list(z_list)
    orig = torch.normal

    def fake(mean, std, size=None, **kw):
        z = queue.pop(0)
        assert tuple(z.shape) == tuple(size), (z.shape, size)
        return mean + std * z

    torch.normal = fake
    try:
        yield queue
    finally:
        torch.normal = orig


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote %s (%.1f KiB)" % (path, os.path.getsize(path) / 1024))


def build_ref_wavenet(ref, sd, cfg):
    m = ref.wavenet.WaveNet_Speech_Commands(**cfg)
    assert list(m.state_dict().keys()) == list(sd.keys()), "state-dict layout mismatch"
    m.load_state_dict(sd)
    return m.eval()


SLICE_T = list(range(0, 16)) + list(range(2040, 2056)) + list(range(8184, 8200)) + list(range(15984, 16000))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = _refharness.load()
    cfg_json = json.load(open(os.path.join(_refharness.REF_ROOT, "configs", "config.json")))
    wcfg, dcfg = cfg_json["wavenet_config"], cfg_json["diffusion_config"]
    assert wcfg == W.DEFAULT_WAVENET_CONFIG and dcfg == W.DEFAULT_DIFFUSION_CONFIG

    # ---- A1/A2 + SDE tables -------------------------------------------------
    hp = ref.util.calc_diffusion_hyperparams(**dcfg)
    steps = torch.tensor([[0.0], [1.0], [33.0], [199.0]])
    emb = ref.util.calc_diffusion_step_embedding(steps, 128)
    sd_full = W.make_state_dict(1234)
    model = build_ref_wavenet(ref, sd_full, wcfg)
    dw = ref.ddpm.DiffWave(model=model, diffusion_hyperparams=hp, reverse_timestep=2)
    rv = ref.sde.RevVPSDE(model=dw, score_type="guided_diffusion", beta_min=0.0001 * 200, beta_max=0.02 * 200,
                          N=200, audio_shape=(1, 16000))
    save("schedule.npz", Beta=hp["Beta"], Alpha=hp["Alpha"], Alpha_bar=hp["Alpha_bar"], Sigma=hp["Sigma"],
         steps=steps, emb=emb, discrete_betas=rv.discrete_betas, alphas_cumprod=rv.alphas_cumprod,
         sqrt_1m_alphas_cumprod=rv.sqrt_1m_alphas_cumprod)

    # ---- A3-A6: full network, one clip, two steps; intermediates sampled -----
    x1 = W.make_waveforms(1, 16000, seed=0)
    inter = {}

    def hook(n):
        def f(mod, inp, out):
            inter[n] = (out[0].detach().clone(), out[1].detach().clone())
        return f

    hs = [model.residual_layer.residual_blocks[n].register_forward_hook(hook(n)) for n in (0, 11, 35)]
    with torch.no_grad():
        eps_t1 = model((x1, 1 * torch.ones((1, 1))))
        keep = {n: inter[n] for n in inter}
        eps_t33 = model((x1, 33 * torch.ones((1, 1))))
    for h in hs:
        h.remove()
    arrays = dict(fingerprint=W.fingerprint(sd_full), x_checksum=np.float64(x1.double().sum()),
                  eps_t1=eps_t1, eps_t33=eps_t33, slice_t=np.asarray(SLICE_T))
    for n, (h, s) in keep.items():
        arrays["h_%d" % n] = h[:, :, SLICE_T]      # (1,256,64): all channels at the sampled times (t=1 run)
        arrays["skip_%d" % n] = s[:, :, SLICE_T]
    save("wavenet_full.npz", **arrays)

    # ---- A3-A6: reduced depth, ragged length, batch 3 ------------------------
    scfg = dict(wcfg, num_res_layers=6, dilation_cycle=3)
    sd_small = W.make_state_dict(99, scfg)
    m_small = build_ref_wavenet(ref, sd_small, scfg)
    xs = W.make_waveforms(3, 1000, seed=5)
    with torch.no_grad():
        eps_s = m_small((xs, 7 * torch.ones((3, 1))))
    save("wavenet_small.npz", fingerprint=W.fingerprint(sd_small), eps=eps_s,
         cfg=json.dumps(scfg), seed=99, x_seed=5, t=7)

    # ---- A7-A9: DiffWave.forward, t*=2 and t*=3, injected noise ---------------
    x2 = W.make_waveforms(2, 16000, seed=0)
    for t_star in (2, 3):
        z = W.make_noise((t_star, 2, 1, 16000), seed=7)
        dw.reverse_timestep = t_star
        with injected_normal([z[i] for i in range(t_star)]) as q:
            y = dw(x2)
            assert not q
        save("ddpm_t%d.npz" % t_star, purified=y, t_star=t_star, x_seed=0, z_seed=7)
        if t_star == 2:
            y_t2 = y

    # ---- A10: one_shot_denoise at reverse_timestep 34 (sigma=0.25) ------------
    dw.reverse_timestep = 34
    with torch.no_grad():
        y1 = dw.one_shot_denoise(x1)
    save("oneshot_t34.npz", x0_hat=y1, reverse_timestep=34)

    # ---- A12: RevVPSDE.f / .g at the solver times of a t=2 integration --------
    t_sde, T = 2, 200
    t0 = 1 - t_sde / T + (-1e-5)
    xf = W.make_waveforms(1, 16000, seed=3).view(1, -1)
    fs, gs, tcs = [], [], []
    for i in range(t_sde):
        tc = torch.tensor(t0 + i / T, dtype=torch.float32)
        with torch.no_grad():
            fs.append(rv.f(tc, xf))
            gs.append(rv.g(tc, xf)[:, :4].clone())
        tcs.append(float(tc))
    save("sde_fg.npz", f=torch.stack(fs), g=torch.stack(gs), tc=np.asarray(tcs), x_seed=3)

    # ---- A14: torchaudio log-mel ----------------------------------------------
    ta = ref.torchaudio
    mel = ta.transforms.MelSpectrogram(n_fft=2048, hop_length=512, n_mels=32, norm="slaney", pad_mode="constant",
                                       mel_scale="slaney")
    a2db = ta.transforms.AmplitudeToDB(stype="power")
    xm = torch.cat([x2, y_t2], dim=0)
    with torch.no_grad():
        spec = a2db(mel(xm))
    save("mel.npz", logmel=spec, fb=mel.mel_scale.fb)

    # ---- A18: ResNeXt-29 8x64 (consumer) ---------------------------------------
    csd = o_resnext.make_state_dict(4321)
    clf = ref.resnext.CifarResNeXt(nlabels=10, in_channels=1)
    assert list(clf.state_dict().keys()) == list(csd.keys())
    clf.load_state_dict(csd)
    clf.eval()
    with torch.no_grad():
        logits = clf(spec)
    save("resnext.npz", logits=logits)

    # ---- A15: AcousticSystem composition (defender -> transform -> classifier) --
    transform = lambda w: a2db(mel(w))  # noqa: E731
    dw.reverse_timestep = 2
    AS = ref.acoustic_system.AcousticSystem(classifier=clf, transform=transform, defender=dw, defense_type="wave")
    z = W.make_noise((2, 2, 1, 16000), seed=7)
    with torch.no_grad(), injected_normal([z[0], z[1]]):
        as_logits = AS(x2)
        as_logits_nodef = AS(x2, defend=False)
    save("acoustic.npz", logits=as_logits, logits_nodefend=as_logits_nodef)

    # ---- A16: RobustCertificate.smooth_predict, 6 draws in batches of 4 --------
    RC = ref.certified.RobustCertificate(classifier=clf, transform=transform, denoiser=dw)
    sigma, n_draw, bs = 0.25, 6, 4
    zc = W.make_noise((n_draw, 1, 16000), seed=11)
    logit_log = []
    orig_forward = RC.forward

    def logging_forward(x):
        out = orig_forward(x)
        logit_log.append(out.detach().clone())
        return out

    RC.forward = logging_forward
    with injected_normal([zc[0:4], zc[4:6]]):
        counts = RC.smooth_predict(x1[0], num_sampling=n_draw, sigma=sigma, batch_size=bs)
    t_star = dw.reverse_timestep
    save("smooth.npz", counts=counts, logits=torch.cat(logit_log, 0), sigma=sigma, t_star=t_star, z_seed=11,
         batch_size=bs)
    print("t_star(sigma=0.25) =", t_star)


if __name__ == "__main__":
    main()
