"""Throughput sweep (BASELINE configs[2] and [4]): global batch x reverse steps, DDPM or reverse VP-SDE, on 1..8 GPUs.

    python tools/sweep.py --batches 1,8,64,256,1024 --tstars 1,2,3,5,10                         # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
        tools/sweep.py --purifier sde --batches 256 --tstars 5,10,25                                # configs[2]

The global batch is split contiguously over the ranks (clip_offset keeps the Philox noise of a clip independent of
the split); there is no collective on the data path.  Time = max over ranks of the CUDA-event device time.  Per row:
purification only and the full purify -> log-mel -> ResNeXt-29 step, whole-network TFLOP/s (batch x steps x 606.10
GFLOP), and per-kernel roofline fractions from a second, event-bracketed pass on rank 0: residual-layer and
tail kernels against the measured bf16 tensor peak, prologue against the measured HBM bandwidth
(MEASURED_PEAKS.json).  Prints a markdown table; --json also writes one JSON object per row."""

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402
from audiopure_b200.certified_robust import shard_range  # noqa: E402

FLOP_PER_CLIP_EVAL = 606.10e9
LAYER_GFLOP = (2 * 512 * 768 * 16000 + 2 * 256 * 256 * 16000) / 1e9
TAIL_GFLOP = (2 * 256 * 36 * 256 * 16000 + 2 * 256 * 256 * 16000 + 2 * 256 * 16000) / 1e9


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,1024")
    p.add_argument("--tstars", default="1,2,3,5,10")
    p.add_argument("--purifier", default="ddpm", choices=["ddpm", "sde"])
    p.add_argument("--json", default=None)
    p.add_argument("--budget-clip-evals", type=int, default=4000, help="clip-evaluations per timing (sets the repeats)")
    args = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6500.0))

    model = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG)
    model.load_state_dict(S.diffwave_state_dict(1234))
    model = model.to(dev).eval()
    eng = model.engine()
    hp = ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(S.resnext_state_dict(4321))
    clf = ap.FusedResNeXt(clf.to(dev).eval()).to(dev)
    tr = ap.LogMelSpectrogram().to(dev)

    def timed(fn, reps):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    rows = []
    if rank == 0:
        print("%s, %d GPU(s); tensor peak %.0f TFLOP/s, HBM peak %.0f GB/s (MEASURED_PEAKS.json)\n" % (args.purifier, world, tc_peak, hbm_peak))
        print("| GPUs | global batch | steps | purify ms | purify clips/s | network TFLOP/s (all GPUs) | full step ms | full clips/s | "
              "layer kernel ms (frac of tensor peak) | tail kernel ms (frac) | prologue us (frac of HBM) | log-mel us | classifier ms |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for B in [int(b) for b in args.batches.split(",")]:
        lo, hi = shard_range(B, rank, world)
        nb = hi - lo
        x = S.clips(max(nb, 1), 16000, seed=B).to(dev)
        for t in [int(t) for t in args.tstars.split(",")]:
            if args.purifier == "sde":
                sde_args = type("A", (), dict(t=t, sample_step=1, rand_t=False, t_delta=0, use_bm=False, score_type="guided_diffusion"))()
                base = ap.DiffWave(model, hp, reverse_timestep=t, seed=1)
                rev = ap.RevDiffWave(sde_args, device=dev, model=base, seed=1)
                purify = (lambda: rev.audio_editing_sample(x, clip_offset=lo)) if nb else (lambda: None)
            else:
                dwm = ap.DiffWave(model, hp, reverse_timestep=t, seed=1)
                purify = (lambda: dwm(x, clip_offset=lo)) if nb else (lambda: None)

            def full():
                if nb:
                    return clf(tr(purify())).max(1)[1]

            per_rank = max(1, (B + world - 1) // world)
            reps = max(1, min(10, args.budget_clip_evals // (per_rank * t)))
            with torch.no_grad():
                ms_p = timed(purify, reps)
                ms_f = timed(full, reps)
                # second pass on this rank: per-kernel device time
                eng = model.engine()
                eng.profile(True)
                eng.profile_read()
                purify()
                torch.cuda.synchronize()
                prof = eng.profile_read()
                eng.profile(False)
                y = purify() if nb else None
                ms_mel = ms_clf = 0.0
                if nb:
                    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    tr(y)
                    clf(tr(y))
                    e[0].record()
                    spec = tr(y)
                    e[1].record()
                    clf(spec)
                    e[2].record()
                    torch.cuda.synchronize()
                    ms_mel, ms_clf = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
            if rank == 0:
                cpl = min(nb, model.max_chunk)  # clips per launch of the first (full) chunk; ragged chunks lower the mean
                layer_ms = prof["layer"][0] / max(prof["layer"][1], 1)
                tail_ms = prof["tail"][0] / max(prof["tail"][1], 1)
                pro_ms = prof["prologue"][0] / max(prof["prologue"][1], 1)
                clips_per_launch = nb * t / max(prof["tail"][1], 1)  # mean over chunks
                lf = clips_per_launch * LAYER_GFLOP / layer_ms / tc_peak if layer_ms else 0.0
                tf = clips_per_launch * TAIL_GFLOP / tail_ms / tc_peak if tail_ms else 0.0
                pf = clips_per_launch * 16000 * (4 + 512) / pro_ms / 1e6 / hbm_peak if pro_ms else 0.0
                row = {"gpus": world, "purifier": args.purifier, "batch": B, "steps": t, "purify_ms": ms_p,
                       "purify_clips_per_s": B / ms_p * 1e3, "network_tflops": B * t * FLOP_PER_CLIP_EVAL / (ms_p * 1e-3) / 1e12,
                       "full_ms": ms_f, "full_clips_per_s": B / ms_f * 1e3, "layer_ms": layer_ms, "layer_frac": lf,
                       "tail_ms": tail_ms, "tail_frac": tf, "prologue_us": pro_ms * 1e3, "prologue_hbm_frac": pf,
                       "logmel_us": ms_mel * 1e3, "classifier_ms": ms_clf, "clips_per_launch": clips_per_launch,
                       "first_chunk": cpl}
                rows.append(row)
                print("| %d | %d | %d | %.2f | %.1f | %.0f | %.2f | %.1f | %.3f (%.2f) | %.3f (%.2f) | %.1f (%.2f) | %.1f | %.2f |" % (
                    world, B, t, ms_p, row["purify_clips_per_s"], row["network_tflops"], ms_f, row["full_clips_per_s"],
                    layer_ms, lf, tail_ms, tf, pro_ms * 1e3, pf, ms_mel * 1e3, ms_clf), flush=True)
    if rank == 0 and args.json:
        with open(args.json, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
