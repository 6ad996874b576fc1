"""Launches the HBM-bound kernels of the path once each at bench size (64 clips) and at 1024 clips, for ncu
(tools/ncu_run.sh), and prints their CUDA-event times (mean of 20 back-to-back launches after warm-up)."""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import _lib, synthetic as S  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    lib = _lib.load()
    tr = ap.LogMelSpectrogram().cuda()
    model = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG)
    model.load_state_dict(S.diffwave_state_dict(1234))
    eng = model.cuda().eval().engine()
    print("| kernel | clips | us per launch | algorithmic MB | GB/s |")
    print("|---|---|---|---|---|")
    for B in (64, 1024):
        x = S.clips(64, 16000, seed=1).cuda().repeat(B // 64, 1, 1).contiguous()
        us = timed(lambda: tr(x))
        mb = B * 68096 / 1e6
        print("| logmel_kernel | %d | %.1f | %.2f | %.0f |" % (B, us, mb, mb / us * 1e3))
        out = torch.empty(B, 1, 16000, device="cuda")
        us = timed(lambda: _lib.check(lib.ap_smooth_inputs(x[0].data_ptr(), 16000, B, 0.25, 0.97, None, 1, 0, 0,
                                                           out.data_ptr(), _lib.stream_ptr())))
        mb = B * 64000 / 1e6
        print("| smooth_inputs_kernel (Philox) | %d | %.1f | %.2f | %.0f |" % (B, us, mb, mb / us * 1e3))
        logits = torch.randn(B, 10, device="cuda")
        counts = torch.zeros(10, dtype=torch.int64, device="cuda")
        us = timed(lambda: _lib.check(lib.ap_vote_counts(logits.data_ptr(), B, 10, counts.data_ptr(), _lib.stream_ptr())))
        print("| vote_counts_kernel | %d | %.1f | %.4f | - |" % (B, us, B * 40 / 1e6))
    x = S.clips(64, 16000, seed=1).cuda()
    eng.profile(True)
    eng.profile_read()
    for _ in range(5):
        eng.eps(x, 1)
    torch.cuda.synchronize()
    prof = eng.profile_read()
    us = prof["prologue"][0] / prof["prologue"][1] * 1e3
    mb = 64 * 16000 * 516 / 1e6
    print("| prologue_kernel | 64 | %.1f | %.1f | %.0f |" % (us, mb, mb / us * 1e3))


if __name__ == "__main__":
    main()
