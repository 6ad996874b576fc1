"""Bring-up diagnostics for the GPU box: each stage runs in its own process (a trapped kernel kills the
CUDA context) and prints where parity breaks.  Test infrastructure (it checks against oracle/), hence under tests/.  Usage: python tests/gpu_diag.py [stage ...]"""

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGES = ["gemm", "small", "full", "purify", "mel", "b64"]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def stage_gemm():
    import torch
    from audiopure_b200 import _lib
    lib = _lib.load()
    for K in (64, 128, 768):
        g = torch.Generator().manual_seed(K)
        a = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
        b = torch.randn(256, K, generator=g).to(torch.bfloat16).cuda()
        d = torch.zeros(128, 256, device="cuda")
        _lib.check(lib.ap_debug_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), K, None))
        torch.cuda.synchronize()
        want = a.float().cpu() @ b.float().cpu().t()
        e = rel(d.cpu(), want)
        print("gemm K=%d rel=%.3e" % (K, e))
        if e > 1e-4:
            dd = (d.cpu() - want).abs()
            print("  err by row%8:", [round(float(dd[r::8].mean()), 3) for r in range(8)])
            print("  err by col block of 32:", [round(float(dd[:, c:c + 32].mean()), 3) for c in range(0, 256, 32)])
            print("  d[0,:8]", d[0, :8].tolist(), "want", want[0, :8].tolist())


def _small():
    import torch
    import audiopure_b200 as ap
    from oracle import weights as W
    cfg = dict(W.DEFAULT_WAVENET_CONFIG, num_res_layers=6, dilation_cycle=3)
    m = ap.WaveNet_Speech_Commands(**cfg)
    m.load_state_dict(W.make_state_dict(99, cfg))
    return m.cuda().eval(), cfg


def stage_small():
    import torch
    from oracle import weights as W, wavenet as o_wavenet
    from tests.emulate import emulate_eps
    m, cfg = _small()
    for (B, L) in ((1, 128), (2, 1000), (3, 2048)):
        x = W.make_waveforms(B, L, seed=11)
        eng = m.engine()
        got = eng.eps(x.cuda(), 3)
        torch.cuda.synchronize()
        packed = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in eng.packed.items()}
        emu, inter = emulate_eps(packed, x, 3, 6, 3, quantize=True, return_inter=True)
        h_bytes = B * L * 512
        ws = eng.workspace(B, L)
        gate = ws[2 * h_bytes: 8 * h_bytes].view(torch.bfloat16).view(6, B, L, 256).float().cpu()
        print("small B=%d L=%d eps rel vs emulation=%.3e" % (B, L, rel(got.cpu(), emu)))
        for n in range(6):
            e = rel(gate[n], inter["gate"][n])
            print("  layer %d gate rel=%.3e" % (n, e))
            if e > 1e-2:
                dd = (gate[n] - inter["gate"][n]).abs()
                print("    err by row%8:", [round(float(dd[:, r::8].mean()), 4) for r in range(8)])
                print("    err by 32-ch block:", [round(float(dd[..., c:c + 32].mean()), 4) for c in range(0, 256, 32)])
                print("    err by 128-row tile:", [round(float(dd[:, r:r + 128].mean()), 4) for r in range(0, L, 128)])
                break
        sd = W.make_state_dict(99, cfg)
        print("  eps rel vs oracle=%.3e" % rel(got.cpu(), o_wavenet.eps_theta(sd, x, 3, cfg)))


def stage_full():
    import numpy as np
    import torch
    import audiopure_b200 as ap
    from oracle import weights as W
    m = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    m.load_state_dict(W.make_state_dict(1234))
    m = m.cuda().eval()
    g = np.load(os.path.join(ROOT, "tests/golden/wavenet_full.npz"))
    x = W.make_waveforms(1, 16000, seed=0)
    for t in (1, 33):
        got = m.engine().eps(x.cuda(), t)
        torch.cuda.synchronize()
        print("full t=%d eps rel vs reference=%.3e" % (t, rel(got.cpu(), torch.from_numpy(g["eps_t%d" % t]))))


def stage_purify():
    import numpy as np
    import torch
    import audiopure_b200 as ap
    from oracle import weights as W
    m = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    m.load_state_dict(W.make_state_dict(1234))
    m = m.cuda().eval()
    hp = ap.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)
    for t_star in (2, 3):
        g = np.load(os.path.join(ROOT, "tests/golden/ddpm_t%d.npz" % t_star))
        x = W.make_waveforms(2, 16000, seed=0)
        z = W.make_noise((t_star, 2, 1, 16000), seed=7)
        y = ap.DiffWave(m, hp, reverse_timestep=t_star)(x.cuda(), z=z)
        torch.cuda.synchronize()
        print("ddpm t*=%d purified rel vs reference=%.3e" % (t_star, rel(y.cpu(), torch.from_numpy(g["purified"]))))


def stage_mel():
    import numpy as np
    import torch
    import audiopure_b200 as ap
    from oracle import weights as W
    g = np.load(os.path.join(ROOT, "tests/golden/mel.npz"))
    x = torch.cat([W.make_waveforms(2, 16000, seed=0),
                   torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/ddpm_t2.npz"))["purified"])], 0)
    got = ap.LogMelSpectrogram().cuda()(x.cuda())
    torch.cuda.synchronize()
    d = (got.cpu() - torch.from_numpy(g["logmel"])).abs()
    print("mel max-abs dB diff=%.3e mean=%.3e" % (float(d.max()), float(d.mean())))


def stage_b64():
    import torch
    import audiopure_b200 as ap
    from oracle import weights as W
    m = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    m.load_state_dict(W.make_state_dict(1234))
    m = m.cuda().eval()
    x = W.make_waveforms(64, 16000, seed=2).cuda()
    eng = m.engine()
    eng.eps(x, 1)
    torch.cuda.synchronize()
    eng.profile(True)
    t0 = time.time()
    y = eng.eps(x, 1)
    torch.cuda.synchronize()
    dt = time.time() - t0
    prof = eng.profile_read()
    eng.profile(False)
    print("B=64 eps wall %.1f ms; finite=%s" % (dt * 1e3, bool(torch.isfinite(y).all())))
    for k, (ms, n) in prof.items():
        print("  %s: %d launches, %.3f ms total, %.3f ms avg" % (k, n, ms, ms / max(n, 1)))
    lms = prof["layer"][0] / prof["layer"][1]
    print("  layer kernel: %.1f TFLOP/s (14.68 GFLOP/clip)" % (64 * 14.68e9 / (lms * 1e-3) / 1e12))
    tms = prof["tail"][0]
    print("  tail kernel: %.1f TFLOP/s (77.6 GFLOP/clip)" % (64 * 77.6e9 / (tms * 1e-3) / 1e12))
    for B in (64, 4, 1):
        xb = x[:B]
        eng.eps(xb, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if B == 64 else 20
        e0.record()
        for _ in range(reps):
            eng.eps(xb, 1)
        e1.record()
        torch.cuda.synchronize()
        print("  unprofiled eval B=%d: %.3f ms" % (B, e0.elapsed_time(e1) / reps))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        globals()["stage_" + sys.argv[2]]()
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        print("=== stage %s" % s, flush=True)
        t0 = time.time()
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], capture_output=True, text=True,
                           timeout=900)
        out = (p.stdout + p.stderr).strip().splitlines()
        out = [l for l in out if "Warning" not in l and "warn" not in l]
        print("\n".join(out[-60:]))
        print("=== stage %s rc=%d %.1fs" % (s, p.returncode, time.time() - t0), flush=True)
