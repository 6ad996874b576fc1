"""Small end-to-end case for compute-sanitizer (memcheck / synccheck) on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitizer_case.py [bf16|tf32]
Reduced-depth network, ragged length, DDPM + SDE + one-shot + log-mel (+ backward; 70 clips so that the persistent
log-mel CTAs loop over several frame pairs) + certification work list (smoothing inputs, votes, ragged batch) + NES."""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

cfg = dict(S.DEFAULT_WAVENET_CONFIG, num_res_layers=3, dilation_cycle=3)
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
if precision == "small":
    # only the non-tensor-core kernels (racecheck stops listing after 100 hazards, and the tcgen05.alloc hand-off of
    # the tensor-core kernels -- a known false positive, profiles/r01_sanitizer.md -- fills that quota)
    from audiopure_b200 import _lib
    from audiopure_b200.blackbox import EOT, NES

    lib = _lib.load()
    tr = ap.LogMelSpectrogram().cuda()
    big = tr(S.clips(70, 16000, seed=4).cuda())
    odd = tr(S.clips(2, 1001, seed=5).cuda())
    xg = S.clips(2, 5000, seed=7).cuda().requires_grad_(True)
    tr(xg).sum().backward()
    x = S.clips(3, 16000, seed=8).cuda().reshape(3, 16000)
    out = torch.empty(10, 1, 16000, device="cuda")
    _lib.check(lib.ap_smooth_inputs_batch(x.data_ptr(), 16000, 10, 7, 9, 0, 0.25, 0.97, None, 1, 0, out.data_ptr(), _lib.stream_ptr()))
    xo = S.clips(2, 1001, seed=9).cuda().reshape(2, 1001)
    out2 = torch.empty(5, 1, 1001, device="cuda")
    _lib.check(lib.ap_smooth_inputs_batch(xo.data_ptr(), 1001, 5, 3, 4, 0, 0.25, 0.97, None, 1, 0, out2.data_ptr(), _lib.stream_ptr()))
    counts = torch.zeros(2, 3, 10, dtype=torch.int64, device="cuda")
    logits = torch.randn(10, 10, device="cuda")
    _lib.check(lib.ap_vote_counts_batch(logits.data_ptr(), 10, 10, 7, 9, 4, 3, counts.data_ptr(), _lib.stream_ptr()))
    toy = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(1001, 10)).cuda()
    nes = NES(8, 4, 0.01, EOT(toy, torch.nn.CrossEntropyLoss(reduction="none"), 2, 1, use_grad=False), seed=1)
    g = nes(S.clips(2, 1001, seed=6).cuda(), torch.tensor([1, 2]))[1]
    torch.cuda.synchronize()
    assert all(torch.isfinite(t).all() for t in (big, odd, xg.grad, out, out2, g)) and int(counts.sum()) == 10
    print("sanitizer case done small", counts.sum((1, 2)).tolist())
    sys.exit(0)
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
xs = S.clips(2, 16000, seed=3).cuda()
yp, rad = rc.certify(xs, torch.zeros(2, dtype=torch.long, device="cuda"), n_0=3, n=7, batch_size=4)
big = tr(S.clips(70, 16000, seed=4).cuda())     # 1120 frame pairs > 1036 resident CTAs
odd = tr(S.clips(2, 1001, seed=5).cuda())       # ragged length: bounds-checked loads, odd frame count
from audiopure_b200.blackbox import EOT, NES  # noqa: E402
toy = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(1001, 10)).cuda()
nes = NES(8, 4, 0.01, EOT(toy, torch.nn.CrossEntropyLoss(reduction="none"), 2, 1, use_grad=False), seed=1)
g = nes(S.clips(2, 1001, seed=6).cuda(), torch.tensor([1, 2]))[1]
torch.cuda.synchronize()
assert all(torch.isfinite(t).all() for t in (y, e, o, r, xg.grad, big, odd, g)) and int(c.sum()) == 6
assert rc.last_counts[0].sum(1).tolist() == [3, 3] and rc.last_counts[1].sum(1).tolist() == [7, 7]
print("sanitizer case done", precision, c.tolist(), yp.tolist())
