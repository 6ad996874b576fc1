"""Multi-GPU checks (run under torchrun on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

1. sharded smoothing: every rank takes a slice of the draw indices, vote counts are summed with
   ap_allreduce_counts (NCCL through the C ABI) and must equal the unsharded counts bit for bit;
2. sharded batch purification with Philox noise equals the unsharded result bit for bit (clip_offset).
"""

import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import audiopure_b200 as ap  # noqa: E402
from audiopure_b200.certified_robust import NcclCountsAllReduce, shard_range  # noqa: E402
from audiopure_b200 import synthetic as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(S.DEFAULT_WAVENET_CONFIG, num_res_layers=6, dilation_cycle=3)
    model = ap.WaveNet_Speech_Commands(**cfg)
    model.load_state_dict(S.diffwave_state_dict(99, cfg))
    model = model.cuda().eval()
    hp = ap.calc_diffusion_hyperparams(**S.DEFAULT_DIFFUSION_CONFIG)
    dw = ap.DiffWave(model, hp, reverse_timestep=2)
    clf = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(S.resnext_state_dict(4321))
    clf = clf.cuda().eval()
    tr = ap.LogMelSpectrogram().cuda()
    x = S.clips(1, 16000, seed=4)[0].cuda()

    n = 203
    allreduce = NcclCountsAllReduce(rank, world)
    lo, hi = shard_range(n, rank, world)
    sharded = ap.RobustCertificate(clf, tr, dw, seed=3, rank=rank, world_size=world, allreduce=allreduce)
    counts = sharded.smooth_predict(x, n, 0.25, batch_size=16)   # the same batch size on every rank
    whole = ap.RobustCertificate(clf, tr, dw, seed=3).smooth_predict(x, n, 0.25, batch_size=n)
    assert int(counts.sum()) == n, counts
    ok1 = torch.equal(counts, whole)
    print("rank %d: sharded counts %s %s unsharded %s" % (rank, counts.tolist(), "==" if ok1 else "!=", whole.tolist()), flush=True)

    # certify over several clips: the flattened (clip, draw) work list sharded over the ranks, ONE all-reduce
    xs = S.clips(3, 16000, seed=14).cuda()
    ys = torch.zeros(3, dtype=torch.long, device="cuda")
    a = ap.RobustCertificate(clf, tr, dw, seed=3, rank=rank, world_size=world, allreduce=allreduce)
    yp, rad = a.certify(xs, ys, n_0=20, n=143, batch_size=16)
    b = ap.RobustCertificate(clf, tr, dw, seed=3)
    yq, rbd = b.certify(xs, ys, n_0=20, n=143, batch_size=16)
    ok1 = ok1 and torch.equal(a.last_counts[0], b.last_counts[0]) and torch.equal(a.last_counts[1], b.last_counts[1]) \
        and torch.equal(yp, yq) and torch.equal(rad, rbd)
    print("rank %d: sharded certify counts %s, y_pred %s radius %s; equal to unsharded: %s"
          % (rank, a.last_counts[1].tolist(), yp.tolist(), [round(r, 4) for r in rad.tolist()], ok1), flush=True)

    B = 4 * world
    xb = S.clips(B, 2048, seed=8).cuda()
    eng = model.engine()
    full = eng.ddpm_purify(xb, 3, seed=77)
    lo, hi = shard_range(B, rank, world)
    mine = eng.ddpm_purify(xb[lo:hi], 3, seed=77, clip_offset=lo)
    ok2 = torch.equal(mine, full[lo:hi])
    print("rank %d: sharded purify slice [%d,%d) bit-equal: %s" % (rank, lo, hi, ok2), flush=True)
    flag = torch.tensor([int(ok1 and ok2)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU CHECK OK (world=%d)" % world)


if __name__ == "__main__":
    main()
