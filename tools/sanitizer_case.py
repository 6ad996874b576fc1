"""Small end-to-end case for compute-sanitizer (memcheck / synccheck) on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitizer_case.py [bf16|tf32]
Reduced-depth network, ragged length, DDPM + SDE + one-shot + log-mel (+ backward) + smoothing inputs + votes."""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

cfg = dict(S.DEFAULT_WAVENET_CONFIG, num_res_layers=3, dilation_cycle=3)
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
m = ap.WaveNet_Speech_Commands(**cfg, precision=precision)
m.load_state_dict(S.diffwave_state_dict(5, cfg))
m = m.cuda().eval()
hp = ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG)
dw = ap.DiffWave(m, hp, reverse_timestep=2)
x = S.waveforms(3, 1000, seed=1).cuda()
y = dw(x)
e = dw.compute_eps_t(x, 3)
o = dw.one_shot_denoise(x)
args = type("A", (), dict(t=2, sample_step=1, rand_t=False, t_delta=0, use_bm=False, score_type="guided_diffusion"))()
r = ap.RevDiffWave(args, model=dw)(x)
tr = ap.LogMelSpectrogram().cuda()
xg = x.clone().requires_grad_(True)
tr(xg).sum().backward()
clf = ap.FusedResNeXt(ap.CifarResNeXt(nlabels=10, in_channels=1).eval()).cuda()  # exercises ap_bias_act_nhwc_bf16
x16 = S.waveforms(1, 16000, seed=2)[0].cuda()
rc = ap.RobustCertificate(clf, tr, ap.DiffWave(m, hp, reverse_timestep=2), seed=1)
c = rc.smooth_predict(x16, 6, 0.25, batch_size=4)
torch.cuda.synchronize()
assert all(torch.isfinite(t).all() for t in (y, e, o, r, xg.grad)) and int(c.sum()) == 6
print("sanitizer case done", precision, c.tolist())
