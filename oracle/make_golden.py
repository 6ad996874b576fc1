"""Oracle tooling (test infrastructure): generate tests/golden/*.npz from the
UNMODIFIED reference, executed on the CPU in the build container.

    python -m oracle.make_golden [--only nes,schedule,wavenet,ddpm,oneshot,sde,frontend,pipeline,certify] [--out DIR]

writes tests/golden/ (everything but `pipeline` (~25 min) and `certify` (~16 min) takes about two minutes).

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


def _logmel_transform(ref):
    ta = ref.torchaudio
    mel = ta.transforms.MelSpectrogram(n_fft=2048, hop_length=512, n_mels=32, norm="slaney", pad_mode="constant",
                                       mel_scale="slaney")
    a2db = ta.transforms.AmplitudeToDB(stype="power")
    return mel, a2db, (lambda w: a2db(mel(w)))


def _ref_classifier(ref):
    csd = o_resnext.make_state_dict(4321)
    clf = ref.resnext.CifarResNeXt(nlabels=10, in_channels=1)
    assert list(clf.state_dict().keys()) == list(csd.keys())
    clf.load_state_dict(csd)
    return clf.eval()


class Ctx:
    """Reference objects shared by the fixture writers (built lazily, once)."""

    def __init__(self):
        torch.set_num_threads(os.cpu_count())
        self.ref = _refharness.load()
        cfg_json = json.load(open(os.path.join(_refharness.REF_ROOT, "configs", "config.json")))
        self.wcfg, self.dcfg = cfg_json["wavenet_config"], cfg_json["diffusion_config"]
        assert self.wcfg == W.DEFAULT_WAVENET_CONFIG and self.dcfg == W.DEFAULT_DIFFUSION_CONFIG
        self.hp = self.ref.util.calc_diffusion_hyperparams(**self.dcfg)
        self.sd_full = W.make_state_dict(1234)
        self.model = build_ref_wavenet(self.ref, self.sd_full, self.wcfg)
        self.dw = self.ref.ddpm.DiffWave(model=self.model, diffusion_hyperparams=self.hp, reverse_timestep=2)
        self.rv = self.ref.sde.RevVPSDE(model=self.dw, score_type="guided_diffusion", beta_min=0.0001 * 200,
                                        beta_max=0.02 * 200, N=200, audio_shape=(1, 16000))
        self.mel, self.a2db, self.transform = _logmel_transform(self.ref)
        self._clf = None

    @property
    def clf(self):
        if self._clf is None:
            self._clf = _ref_classifier(self.ref)
        return self._clf


def fx_schedule(c):
    # ---- A1/A2 + SDE tables -------------------------------------------------
    steps = torch.tensor([[0.0], [1.0], [33.0], [199.0]])
    emb = c.ref.util.calc_diffusion_step_embedding(steps, 128)
    save("schedule.npz", Beta=c.hp["Beta"], Alpha=c.hp["Alpha"], Alpha_bar=c.hp["Alpha_bar"], Sigma=c.hp["Sigma"],
         steps=steps, emb=emb, discrete_betas=c.rv.discrete_betas, alphas_cumprod=c.rv.alphas_cumprod,
         sqrt_1m_alphas_cumprod=c.rv.sqrt_1m_alphas_cumprod)


def fx_wavenet(c):
    # ---- A3-A6: full network, one clip, two steps; intermediates sampled -----
    model = c.model
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
    arrays = dict(fingerprint=W.fingerprint(c.sd_full), x_checksum=np.float64(x1.double().sum()),
                  eps_t1=eps_t1, eps_t33=eps_t33, slice_t=np.asarray(SLICE_T))
    for n, (h, s) in keep.items():
        arrays["h_%d" % n] = h[:, :, SLICE_T]      # (1,256,64): all channels at the sampled times (t=1 run)
        arrays["skip_%d" % n] = s[:, :, SLICE_T]
    save("wavenet_full.npz", **arrays)

    # ---- A3-A6: reduced depth, ragged length, batch 3 ------------------------
    scfg = dict(c.wcfg, num_res_layers=6, dilation_cycle=3)
    sd_small = W.make_state_dict(99, scfg)
    m_small = build_ref_wavenet(c.ref, sd_small, scfg)
    xs = W.make_waveforms(3, 1000, seed=5)
    with torch.no_grad():
        eps_s = m_small((xs, 7 * torch.ones((3, 1))))
    save("wavenet_small.npz", fingerprint=W.fingerprint(sd_small), eps=eps_s,
         cfg=json.dumps(scfg), seed=99, x_seed=5, t=7)


def fx_ddpm(c):
    # ---- A7-A9: DiffWave.forward, t*=2 and t*=3, injected noise ---------------
    x2 = W.make_waveforms(2, 16000, seed=0)
    for t_star in (2, 3):
        z = W.make_noise((t_star, 2, 1, 16000), seed=7)
        c.dw.reverse_timestep = t_star
        with injected_normal([z[i] for i in range(t_star)]) as q:
            y = c.dw(x2)
            assert not q
        save("ddpm_t%d.npz" % t_star, purified=y, t_star=t_star, x_seed=0, z_seed=7)


def fx_oneshot(c):
    # ---- A10: one_shot_denoise.  t=34 on a clean white-noise clip (round-1 fixture), and at the t* of sigma = 0.1,
    # 0.5, 1.0 (certified_robust.py:102-110 -> 14, 66, 117) on the certifier's own kind of input:
    # sqrt(abar*) (x + sigma z) of a structured clip (certified_robust.py:46-54) ----
    x1 = W.make_waveforms(1, 16000, seed=0)
    c.dw.reverse_timestep = 34
    with torch.no_grad():
        y1 = c.dw.one_shot_denoise(x1)
    save("oneshot_t34.npz", x0_hat=y1, reverse_timestep=34)
    xc = W.make_clips(1, 16000, seed=40)
    zc = W.make_noise((1, 1, 16000), seed=41)
    arrays = {}
    for sigma, t_star in ((0.1, 14), (0.5, 66), (1.0, 117)):
        ab = 1 / (1 + sigma ** 2)
        assert torch.abs(c.hp["Alpha_bar"] - ab).min(0, keepdim=True)[1].item() + 1 == t_star
        c.dw.reverse_timestep = t_star
        with torch.no_grad():
            arrays["x0_hat_t%d" % t_star] = c.dw.one_shot_denoise(ab ** 0.5 * (xc + sigma * zc))
    save("oneshot_sigmas.npz", x_seed=40, z_seed=41, sigmas=np.asarray([0.1, 0.5, 1.0]),
         t_stars=np.asarray([14, 66, 117]), **arrays)


def fx_sde(c):
    # ---- A12: RevVPSDE.f / .g at the solver times of a t=2 integration --------
    rv = c.rv
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

    # ---- A13 at t = 5 (B = 2) and t = 10 (B = 1): diffwave_sde.py:183-205 with the reference's OWN RevVPSDE.f / .g
    # driving a fixed-step Euler-Maruyama loop (torchsde 0.2.5 is absent; its method='euler' is
    # y <- y + f dt + g dW, dW = sqrt(dt) z, restated here -- "parity unpinned" for the stepping only) ----
    for t_sde, B in ((5, 2), (10, 1)):
        x0 = W.make_clips(B, 16000, seed=50 + t_sde)
        z = W.make_noise((t_sde + 1, B, 1, 16000), seed=60 + t_sde)
        betas = rv.discrete_betas.float()
        a = (1 - betas).cumprod(dim=0)
        x = x0 * a[t_sde - 1].sqrt() + z[0] * (1.0 - a[t_sde - 1]).sqrt()     # diffwave_sde.py:185-191
        t0 = 1 - t_sde / T + (-1e-5)
        dt = 1.0 / T
        y = x.view(B, -1)
        with torch.no_grad():
            for i in range(t_sde):
                tc = torch.tensor(t0 + i * dt, dtype=torch.float32)
                y = y + rv.f(tc, y) * dt + rv.g(tc, y) * (z[1 + i].view(B, -1) * dt ** 0.5)
        save("sde_t%d.npz" % t_sde, purified=y.view(B, 1, 16000), t=t_sde, x_seed=50 + t_sde, z_seed=60 + t_sde)


def fx_frontend(c):
    # ---- A14: torchaudio log-mel; A18: ResNeXt-29 8x64 (consumer, calibrated synthetic checkpoint);
    # A15: AcousticSystem composition (defender -> transform -> classifier) --
    x2 = W.make_waveforms(2, 16000, seed=0)
    src = os.path.join(OUT, "ddpm_t2.npz")  # written by the `ddpm` step; fall back to the committed fixture
    if not os.path.exists(src):
        src = os.path.join(ROOT, "tests", "golden", "ddpm_t2.npz")
    y_t2 = torch.from_numpy(np.load(src)["purified"])
    xm = torch.cat([x2, y_t2, W.make_clips(4, 16000, seed=20)], dim=0)
    with torch.no_grad():
        spec = c.transform(xm)
    save("mel.npz", logmel=spec, fb=c.mel.mel_scale.fb)
    with torch.no_grad():
        logits = c.clf(spec)
    save("resnext.npz", logits=logits)
    xa = W.make_clips(2, 16000, seed=21)
    c.dw.reverse_timestep = 2
    AS = c.ref.acoustic_system.AcousticSystem(classifier=c.clf, transform=c.transform, defender=c.dw, defense_type="wave")
    z = W.make_noise((2, 2, 1, 16000), seed=7)
    with torch.no_grad(), injected_normal([z[0], z[1]]):
        as_logits = AS(xa)
        as_logits_nodef = AS(xa, defend=False)
    save("acoustic.npz", logits=as_logits, logits_nodefend=as_logits_nodef, x_seed=21, z_seed=7)


PIPE_N, PIPE_CHUNK, PIPE_STRIDE = 256, 16, 125


def fx_pipeline(c):
    # ---- BASELINE configs[0]/[1] workload through the reference's AcousticSystem on 256 structured clips
    # (DDPM t*=2 -> log-mel -> ResNeXt), injected noise; ~25 min of CPU.  Stored: logits, and the purified
    # waveform at every 125th sample plus the first and last 8 samples of every clip. ----
    x = W.make_clips(PIPE_N, 16000, seed=31)
    z = W.make_noise((2, PIPE_N, 1, 16000), seed=32)
    c.dw.reverse_timestep = 2
    keep = {}

    class Tap(torch.nn.Module):
        def forward(self, w):
            keep["y"] = c.dw(w)
            return keep["y"]

    AS = c.ref.acoustic_system.AcousticSystem(classifier=c.clf, transform=c.transform, defender=Tap(), defense_type="wave")
    idx = np.unique(np.concatenate([np.arange(0, 16000, PIPE_STRIDE), np.arange(8), np.arange(15992, 16000)]))
    logits, samples = [], []
    for s in range(0, PIPE_N, PIPE_CHUNK):
        e = s + PIPE_CHUNK
        with torch.no_grad(), injected_normal([z[0, s:e], z[1, s:e]]) as q:
            logits.append(AS(x[s:e]))
            assert not q
        samples.append(keep["y"][:, 0, idx].clone())
        print("pipeline: %d / %d clips" % (e, PIPE_N), flush=True)
    save("pipeline256.npz", logits=torch.cat(logits), purified_samples=torch.cat(samples), sample_idx=idx,
         x_seed=31, z_seed=32, t_star=2)


CERT_CLIPS, CERT_N0, CERT_N, CERT_BS, CERT_SIGMA, CERT_ALPHA = (1, 6), 32, 128, 64, 0.25, 0.001


def fx_certify(c):
    # ---- A16/A17: RobustCertificate.certify on two structured clips (one certifies, one abstains), n_0 = 32,
    # n = 128, sigma = 0.25, batches of 64, injected noise; every draw's logits logged.  ~16 min of CPU.
    # statsmodels is absent: proportion_confint(method='beta') is supplied as Clopper-Pearson through scipy
    # (oracle.certify.lower_conf_bound) -- parity unpinned for that one call. ----
    from oracle import certify as o_certify
    import sys as _sys

    def proportion_confint(count, nobs, alpha=0.05, method="normal"):
        assert method == "beta"
        return o_certify.lower_conf_bound(int(count), int(nobs), alpha / 2), None

    _sys.modules["statsmodels.stats.proportion"].proportion_confint = proportion_confint
    c.ref.certified.proportion_confint = proportion_confint
    RC = c.ref.certified.RobustCertificate(classifier=c.clf, transform=c.transform, denoiser=c.dw)
    xs = W.make_clips(8, 16000, seed=31)[list(CERT_CLIPS)]
    zc = W.make_noise((len(CERT_CLIPS), CERT_N0 + CERT_N, 1, 16000), seed=11)
    logit_log, count_log = [], []
    orig_forward, orig_sp = RC.forward, RC.smooth_predict

    def logging_forward(x):
        out = orig_forward(x)
        logit_log.append(out.detach().clone())
        return out

    def logging_sp(*a, **k):
        out = orig_sp(*a, **k)
        count_log.append(out.detach().clone())
        return out

    RC.forward, RC.smooth_predict = logging_forward, logging_sp
    queue = []
    for i in range(len(CERT_CLIPS)):
        queue.append(zc[i, :CERT_N0])
        for s in range(CERT_N0, CERT_N0 + CERT_N, CERT_BS):
            queue.append(zc[i, s:s + CERT_BS])
    y = torch.zeros(len(CERT_CLIPS), dtype=torch.long)
    with injected_normal(queue) as q:
        y_pred, radius = RC.certify(xs, y, sigma=CERT_SIGMA, n_0=CERT_N0, n=CERT_N, alpha=CERT_ALPHA, batch_size=CERT_BS)
        assert not q
    logits = torch.cat(logit_log, 0).reshape(len(CERT_CLIPS), CERT_N0 + CERT_N, -1)
    counts_0 = torch.stack(count_log[0::2])
    counts = torch.stack(count_log[1::2])
    print("certify: counts_0", counts_0.tolist(), "counts", counts.tolist(), "y_pred", y_pred.tolist(), "radius", radius.tolist())
    save("certify.npz", logits=logits, counts_0=counts_0, counts=counts, y_pred=y_pred, radius=radius,
         clips=np.asarray(CERT_CLIPS), x_seed=31, z_seed=11, sigma=CERT_SIGMA, n_0=CERT_N0, n=CERT_N, alpha=CERT_ALPHA,
         batch_size=CERT_BS, t_star=c.dw.reverse_timestep)


@contextlib.contextmanager
def injected_randn(z_list):
    queue = list(z_list)
    orig = torch.randn

    def fake(size, **kw):
        z = queue.pop(0)
        assert tuple(z.shape) == tuple(size), (z.shape, size)
        return z.clone()

    torch.randn = fake
    try:
        yield queue
    finally:
        torch.randn = orig


def fx_nes(c):
    # ---- section 8(f)3: the reference's own NES(EOT(model)) classes (robustness_eval/_NES.py, _EOT.py) on a toy
    # deterministic model, injected noise: 3 audios x 256 samples, 8 samples per draw in batches of 4, EOT 2x1 ----
    import importlib

    from oracle import blackbox as o_bb

    ref_nes = importlib.import_module("robustness_eval._NES")
    ref_eot = importlib.import_module("robustness_eval._EOT")
    n_audios, N, spd, S, sigma = 3, 256, 8, 4, 0.01
    x = W.make_clips(n_audios, N, seed=70)
    y = torch.tensor([1, 4, 7])
    z = W.make_noise((spd // S, n_audios, S // 2, 1, N), seed=71)
    model = o_bb.toy_model(N)
    loss = torch.nn.CrossEntropyLoss(reduction="none")           # _utils.py:116-117 (task 'SCR')
    eot = ref_eot.EOT(model, loss, EOT_size=2, EOT_batch_size=1, use_grad=False)
    nes = ref_nes.NES(spd, S, sigma, eot)
    with torch.no_grad(), injected_randn([z[i] for i in range(spd // S)]) as q:
        mean_loss, grad, adver_loss, adver_score, predict = nes(x, y)
        assert not q
    save("nes.npz", mean_loss=mean_loss, grad=grad, adver_loss=adver_loss, adver_score=adver_score,
         predict=np.asarray(predict), x_seed=70, z_seed=71, y=y, samples_per_draw=spd, samples_per_draw_batch=S,
         sigma=sigma, EOT_size=2, EOT_batch_size=1, N=N, model_seed=77)


FIXTURES = {"nes": fx_nes, "schedule": fx_schedule, "wavenet": fx_wavenet, "ddpm": fx_ddpm, "oneshot": fx_oneshot, "sde": fx_sde,
            "frontend": fx_frontend, "pipeline": fx_pipeline, "certify": fx_certify}


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated subset of: " + ", ".join(FIXTURES))
    ap.add_argument("--out", default=None, help="output directory (default tests/golden)")
    args = ap.parse_args()
    global OUT
    if args.out:
        OUT = args.out
    os.makedirs(OUT, exist_ok=True)
    names = [n for n in args.only.split(",") if n] or list(FIXTURES)
    c = Ctx()
    for n in names:
        print("== %s" % n, flush=True)
        FIXTURES[n](c)


if __name__ == "__main__":
    main()
