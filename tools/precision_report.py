"""bf16 vs tf32 mode on the GPU box: errors against the reference fixtures and kernel timings.

    python tools/precision_report.py > gpurun_out/precision.md

Uses tests/golden/*.npz (generated from the unmodified reference, oracle/make_golden.py) as the fp32 truth.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def main():
    hp = ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG)
    g_net = np.load(os.path.join(GOLD, "wavenet_full.npz"))
    g_os = np.load(os.path.join(GOLD, "oneshot_t34.npz"))
    rows = []
    timing = []
    for prec in ("bf16", "tf32"):
        m = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG, precision=prec)
        m.load_state_dict(S.diffwave_state_dict(1234))
        m = m.cuda().eval()
        eng = m.engine()
        x1 = S.waveforms(1, 16000, seed=0).cuda()
        r = {"precision": prec, "round_bias": hex(eng.precision()[1])}
        for t in (1, 33):
            r["eps_t%d" % t] = rel(eng.eps(x1, t), g_net["eps_t%d" % t])
        for ts in (2, 3):
            g = np.load(os.path.join(GOLD, "ddpm_t%d.npz" % ts))
            x = S.waveforms(2, 16000, seed=int(g["x_seed"])).cuda()
            z = S.noise((ts, 2, 1, 16000), seed=int(g["z_seed"]))
            r["ddpm_t%d" % ts] = rel(ap.DiffWave(m, hp, reverse_timestep=ts)(x, z=z), g["purified"])
        dw = ap.DiffWave(m, hp, reverse_timestep=int(g_os["reverse_timestep"]))
        r["oneshot_t34"] = rel(dw.one_shot_denoise(x1), g_os["x0_hat"])
        rows.append(r)

        B = 32
        xb = S.waveforms(B, 16000, seed=1).cuda()
        for _ in range(3):
            eng.eps(xb, 1)
        torch.cuda.synchronize()
        eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            eng.eps(xb, 1)
        e1.record()
        torch.cuda.synchronize()
        prof = eng.profile_read()
        eng.profile(False)
        ms = e0.elapsed_time(e1) / n
        flop = 603.98e9 * B
        timing.append({"precision": prec, "B": B, "eval_ms": ms, "TFLOP/s": flop / ms / 1e9,
                       "layer_ms": prof["layer"][0] / prof["layer"][1], "tail_ms": prof["tail"][0] / prof["tail"][1]})
        del eng, m
        torch.cuda.empty_cache()

    print("# bf16 vs tf32 mode (full 36-layer network, reference fixtures as fp32 truth)\n")
    keys = [k for k in rows[0] if k not in ("precision",)]
    print("| precision | " + " | ".join(keys) + " |")
    print("|---|" + "---|" * len(keys))
    for r in rows:
        print("| %s | " % r["precision"] + " | ".join(("%.3e" % r[k]) if isinstance(r[k], float) else str(r[k]) for k in keys) + " |")
    print("\nrel-L2 of eps (network output) and of the purified waveform; north_star gates: waveform <= 1e-2 (bf16), <= 1e-3 (tf32).\n")
    print("| precision | batch | eval ms | residual-stack TFLOP/s | layer kernel ms | tail kernel ms |")
    print("|---|---|---|---|---|---|")
    for t in timing:
        print("| %s | %d | %.2f | %.0f | %.3f | %.3f |" % (t["precision"], t["B"], t["eval_ms"], t["TFLOP/s"], t["layer_ms"], t["tail_ms"]))


if __name__ == "__main__":
    main()
