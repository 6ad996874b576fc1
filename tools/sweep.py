"""Throughput sweep (BASELINE configs[4]): batch x t* on one GPU, DDPM purification only and the full
purify -> log-mel -> ResNeXt-29 step.  Prints a markdown table (device time, CUDA events)."""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

FLOP_PER_CLIP_EVAL = 606.10e9


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    torch.backends.cudnn.benchmark = True
    model = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG)
    model.load_state_dict(S.diffwave_state_dict(1234))
    model = model.cuda().eval()
    hp = ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(S.resnext_state_dict(4321))
    clf = ap.FusedResNeXt(clf.cuda().eval()).cuda()
    tr = ap.LogMelSpectrogram().cuda()
    batches = [int(b) for b in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,2,4,8,16,32,64,128,256,1024".split(","))]
    tstars = [int(t) for t in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,3,5,10".split(","))]
    print("| batch | t* | purify ms | purify clips/s | network TFLOP/s | full step ms | full clips/s |\n|---|---|---|---|---|---|---|")
    for B in batches:
        x = S.waveforms(B, 16000, seed=B).cuda()
        for t in tstars:
            dw = ap.DiffWave(model, hp, reverse_timestep=t)
            system = ap.AcousticSystem(clf, tr, dw)
            reps = max(1, min(10, int(2000 / (B * t))))
            with torch.no_grad():
                ms_p = timed(lambda: dw(x), reps)
                ms_f = timed(lambda: system(x).max(1)[1], reps)
            print("| %d | %d | %.2f | %.1f | %.0f | %.2f | %.1f |" % (
                B, t, ms_p, B / ms_p * 1e3, B * t * FLOP_PER_CLIP_EVAL / (ms_p * 1e-3) / 1e12, ms_f, B / ms_f * 1e3), flush=True)


if __name__ == "__main__":
    main()
