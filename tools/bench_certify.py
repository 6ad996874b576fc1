"""Randomized-smoothing certification throughput (BASELINE configs[3]): draws sharded over the ranks, only
the int64 vote counts all-reduced (NCCL through the C ABI).

    python tools/bench_certify.py [--draws 10000] [--clips 1]                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29520 tools/bench_certify.py --draws 10000                  # 8 GPUs

Prints one JSON line: smoothing draws/s over all ranks (device time, max over ranks).
"""

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200.certified_robust import NcclCountsAllReduce  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--draws", dest="n", type=int, default=10000)
    p.add_argument("--select-draws", dest="n0", type=int, default=100)
    p.add_argument("--clips", type=int, default=1)
    p.add_argument("--sigma", type=float, default=0.25)
    p.add_argument("--batch", type=int, default=64)
    args = p.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    torch.backends.cudnn.benchmark = False  # same consumer kernels in every process: identical vote counts at every N
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    allreduce = NcclCountsAllReduce(rank, world)

    model = ap.WaveNet_Speech_Commands(**S.DEFAULT_WAVENET_CONFIG)
    model.load_state_dict(S.diffwave_state_dict(1234))
    model = model.cuda().eval()
    dw = ap.DiffWave(model, ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG), reverse_timestep=34)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(S.resnext_state_dict(4321))
    clf = ap.FusedResNeXt(clf.cuda().eval()).cuda()
    RC = ap.RobustCertificate(clf, ap.LogMelSpectrogram().cuda(), dw, seed=5, rank=rank, world_size=world,
                              allreduce=allreduce)
    x = S.clips(args.clips, 16000, seed=3).cuda()
    y = torch.zeros(args.clips, dtype=torch.long, device="cuda")

    # warm-up with exactly the timed call's batch shapes (cudnn.benchmark autotunes once per shape)
    RC.certify(x, y, sigma=args.sigma, n_0=args.n0, n=args.n, batch_size=args.batch, clip_offset=0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y_pred, radius = RC.certify(x, y, sigma=args.sigma, n_0=args.n0, n=args.n, batch_size=args.batch, clip_offset=0)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if rank == 0:
        draws = args.clips * (args.n + args.n0)
        print(json.dumps({"metric": "smoothing draws/sec (one-shot denoise t*=34 + log-mel + ResNeXt-29 per draw)",
                          "value": draws / (float(ms) * 1e-3), "unit": "draws/s", "n_gpus": world,
                          "clips": args.clips, "n": args.n, "n0": args.n0, "sigma": args.sigma,
                          "seconds_per_clip": float(ms) * 1e-3 / args.clips,
                          "y_pred": y_pred.tolist(), "radius": [round(r, 4) for r in radius.tolist()]}))


if __name__ == "__main__":
    main()
