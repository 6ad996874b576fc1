"""Oracle tooling (test infrastructure): calibrate the synthetic ResNeXt-29 checkpoint.

    python -m oracle.calibrate_resnext        # writes audiopure_b200/data/resnext_calib.npz

Why.  No trained Speech-Commands checkpoint is reachable offline.  A raw random init of
``resnext.py:88-111`` (unit batch-norm statistics, zero biases) on log-mel images in dB is a positively
homogeneous map dominated by the image's mean level: it predicts ONE class for every input, so "top-1
agreement" and "vote counts bit-exact" comparisons could not fail.  This script gives the random-init
network what training would have given it as far as those comparisons are concerned:

* every batch-norm's running mean / variance = the statistics of its own input over a calibration batch
  of synthetic log-mels (clean and noise-smoothed structured clips, ``oracle.weights.make_clips``),
  computed layer by layer like one training-mode forward pass (``nn.BatchNorm2d`` with momentum=None);
* the final linear layer is centred on the calibration features and scaled so that the logits have
  unit-order spread: predictions cover all classes and a few percent of clips are near-ties.

Convolution weights, batch-norm affine parameters and the classifier direction still come from the numpy
PCG64 stream of ``oracle.resnext.make_state_dict``; only the data-dependent tensors are stored, because a
forward pass is not bit-reproducible across machines and every machine must build the SAME checkpoint.
"""

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import mel as o_mel, resnext as o_resnext, weights as W  # noqa: E402

OUT = os.path.join(ROOT, "audiopure_b200", "data", "resnext_calib.npz")
N_CLEAN, N_NOISY = 192, 192
LOGIT_STD = 2.5


def calibration_batch():
    clean = W.make_clips(N_CLEAN, 16000, seed=100)
    base = W.make_clips(N_NOISY, 16000, seed=101)
    z = W.make_noise((N_NOISY, 1, 16000), seed=102)
    sig = torch.tensor([0.25, 0.5]).repeat(N_NOISY // 2).reshape(-1, 1, 1)
    return o_mel.log_mel(torch.cat([clean, base + sig * z], 0))


def main():
    torch.set_num_threads(os.cpu_count())
    if os.path.exists(OUT):
        os.unlink(OUT)
    sd = o_resnext.make_state_dict(4321)  # jittered batch-norm affine, statistics still at their init
    x = calibration_batch()
    stats = {}

    def bn(sd_, prefix, v):
        mean = v.mean(dim=(0, 2, 3))
        var = v.var(dim=(0, 2, 3), unbiased=False)
        sd_[prefix + ".running_mean"] = mean
        sd_[prefix + ".running_var"] = var
        stats[prefix + ".running_mean"] = mean.numpy().copy()
        stats[prefix + ".running_var"] = var.numpy().copy()
        return F.batch_norm(v, mean, var, sd_[prefix + ".weight"], sd_[prefix + ".bias"], training=False, eps=1e-5)

    with torch.no_grad():
        feats = o_resnext.features(sd, x, bn=bn)                      # (N, 1024)
        w = sd["classifier.weight"]
        centred = feats - feats.mean(0, keepdim=True)
        raw = centred @ w.t()
        scale = LOGIT_STD / float(raw.std())
        w = w * scale
        b = -(feats.mean(0, keepdim=True) @ w.t())[0]
        sd["classifier.weight"], sd["classifier.bias"] = w, b
        logits = o_resnext.forward(sd, x)
    stats["classifier.weight"] = w.numpy().copy()
    stats["classifier.bias"] = b.numpy().copy()
    pred = logits.argmax(1)
    top2 = logits.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    print("classes hit:", sorted(set(pred.tolist())), "hist:", np.bincount(pred.numpy(), minlength=10).tolist())
    print("logit std %.3f  margins < 0.05: %d / %d" % (float(logits.std()), int((margin < 0.05).sum()), len(margin)))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **stats)
    print("wrote %s (%.1f KiB)" % (OUT, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
