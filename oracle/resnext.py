"""Oracle (test infrastructure): ResNeXt-29 8x64 forward, restated functionally.

Restates ``audio_models/ConvNets_SpeechCommands/models/resnext.py:56-64,134-142``
(inference / ``eval()`` mode: batch-norm uses running statistics) over a
reference-layout state dict.  The classifier is a *consumer* of the hot path
(SURVEY.md section 2 row 8): the product keeps it a cuDNN ``nn.Module``; this
restatement exists so logits / votes can be checked on the CPU.
"""

import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

CARDINALITY = 8
DEPTH = 29
BASE_WIDTH = 64
WIDEN = 4


def _stage_plan():
    stages = [64, 64 * WIDEN, 128 * WIDEN, 256 * WIDEN]
    block_depth = (DEPTH - 2) // 9
    plan = []
    for s, stride in ((1, 1), (2, 2), (3, 2)):
        cin, cout = stages[s - 1], stages[s]
        for j in range(block_depth):
            width_ratio = cout / (WIDEN * 64.0)
            D = CARDINALITY * int(BASE_WIDTH * width_ratio)
            plan.append(("stage_%d.stage_%d_bottleneck_%d" % (s, s, j), cin if j == 0 else cout, cout,
                         stride if j == 0 else 1, D))
    return stages, plan


def _bn(sd, prefix, x):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=1e-5)


def features(sd, x, bn=_bn):
    """resnext.py:134-140 (everything before the linear layer): (B,1,32,32) -> (B,1024).  ``bn`` is replaceable so
    that oracle/calibrate_resnext.py can collect batch statistics with the same dataflow."""
    stages, plan = _stage_plan()
    x = F.relu(bn(sd, "bn_1", F.conv2d(x, sd["conv_1_3x3.weight"], padding=1)))
    for name, cin, cout, stride, D in plan:
        b = F.relu(bn(sd, name + ".bn_reduce", F.conv2d(x, sd[name + ".conv_reduce.weight"])))
        b = F.relu(bn(sd, name + ".bn", F.conv2d(b, sd[name + ".conv_conv.weight"], stride=stride, padding=1,
                                                  groups=CARDINALITY)))
        b = bn(sd, name + ".bn_expand", F.conv2d(b, sd[name + ".conv_expand.weight"]))
        if cin != cout:
            r = bn(sd, name + ".shortcut.shortcut_bn",
                   F.conv2d(x, sd[name + ".shortcut.shortcut_conv.weight"], stride=stride))
        else:
            r = x
        x = F.relu(r + b)
    x = F.avg_pool2d(x, 8, 1)
    return x.view(-1, stages[3])


def forward(sd, x):
    """resnext.py:134-142 over state dict ``sd``; x: (B,1,32,32) -> (B,nlabels)."""
    return F.linear(features(sd, x), sd["classifier.weight"], sd["classifier.bias"])


CALIB_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audiopure_b200", "data",
                          "resnext_calib.npz")


def make_state_dict(seed=4321, nlabels=10, in_channels=1, calibrated=True):
    """Deterministic synthetic checkpoint in the reference layout, from a numpy PCG64 stream.

    ``calibrated=False``: the constructor's own init (resnext.py:88-111: kaiming-normal fan_out convs, unit
    batch-norm, zero biases) -- it predicts one class for every log-mel input.  ``calibrated=True`` (default, seed
    4321 only): batch-norm affine parameters are jittered (scale U(0.7,1.3), shift N(0,0.1)) and the
    data-dependent tensors -- batch-norm running statistics and the centred, rescaled linear layer -- are read
    from ``audiopure_b200/data/resnext_calib.npz`` (written once by oracle/calibrate_resnext.py), so predictions
    on synthetic clips span the classes and some are near-ties."""
    if calibrated:
        assert seed == 4321 and nlabels == 10 and in_channels == 1, "the calibration file is for seed 4321"
        sd = make_state_dict(seed, nlabels, in_channels, calibrated=False)
        rng2 = np.random.Generator(np.random.PCG64(seed + 1))
        for k in [k for k in sd if k.endswith("running_mean")]:
            p = k[:-len(".running_mean")]
            c = sd[k].numel()
            sd[p + ".weight"] = torch.from_numpy(rng2.uniform(0.7, 1.3, size=c).astype(np.float32))
            sd[p + ".bias"] = torch.from_numpy(rng2.normal(0.0, 0.1, size=c).astype(np.float32))
        if os.path.exists(CALIB_FILE):  # absent only while oracle/calibrate_resnext.py is creating it
            with np.load(CALIB_FILE) as cal:
                for k in cal.files:
                    assert tuple(cal[k].shape) == tuple(sd[k].shape), k
                    sd[k] = torch.from_numpy(cal[k].copy())
        return sd
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = OrderedDict()

    def conv(name, cout, cin, k):
        std = np.sqrt(2.0 / (cout * k * k))
        sd[name + ".weight"] = torch.from_numpy(rng.normal(0, std, size=(cout, cin, k, k)).astype(np.float32))

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    stages, plan = _stage_plan()
    conv("conv_1_3x3", 64, in_channels, 3)
    bn("bn_1", 64)
    for name, cin, cout, stride, D in plan:
        conv(name + ".conv_reduce", D, cin, 1)
        bn(name + ".bn_reduce", D)
        conv(name + ".conv_conv", D, D // CARDINALITY, 3)
        bn(name + ".bn", D)
        conv(name + ".conv_expand", cout, D, 1)
        bn(name + ".bn_expand", cout)
        if cin != cout:
            conv(name + ".shortcut.shortcut_conv", cout, cin, 1)
            bn(name + ".shortcut.shortcut_bn", cout)
    sd["classifier.weight"] = torch.from_numpy(
        rng.normal(0, np.sqrt(2.0 / stages[3]), size=(nlabels, stages[3])).astype(np.float32))
    sd["classifier.bias"] = torch.zeros(nlabels)
    return sd
