"""ResNeXt-29 8x64 Speech-Commands classifier -- the *consumer* of the purification path.

Not accelerated here (< 1 % of the path's FLOPs; it stays a cuDNN
``nn.Module``), but the package needs the architecture to compose and benchmark the full
purify -> log-mel -> classify step without the reference checkout.  Same layer names as
``audio_models/ConvNets_SpeechCommands/models/resnext.py`` so its checkpoints' state dicts load.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

__all__ = ["CifarResNeXt", "FusedResNeXt"]


class ResNeXtBottleneck(nn.Module):
    """resnext.py:23-64 (type-C bottleneck: 1x1 reduce, grouped 3x3, 1x1 expand, projection shortcut)."""

    def __init__(self, in_channels, out_channels, stride, cardinality, base_width, widen_factor):
        super().__init__()
        width_ratio = out_channels / (widen_factor * 64.0)
        D = cardinality * int(base_width * width_ratio)
        self.conv_reduce = nn.Conv2d(in_channels, D, 1, 1, 0, bias=False)
        self.bn_reduce = nn.BatchNorm2d(D)
        self.conv_conv = nn.Conv2d(D, D, 3, stride, 1, groups=cardinality, bias=False)
        self.bn = nn.BatchNorm2d(D)
        self.conv_expand = nn.Conv2d(D, out_channels, 1, 1, 0, bias=False)
        self.bn_expand = nn.BatchNorm2d(out_channels)
        self.shortcut = nn.Sequential()
        if in_channels != out_channels:
            self.shortcut.add_module("shortcut_conv", nn.Conv2d(in_channels, out_channels, 1, stride, 0, bias=False))
            self.shortcut.add_module("shortcut_bn", nn.BatchNorm2d(out_channels))

    def forward(self, x):
        y = F.relu(self.bn_reduce(self.conv_reduce(x)), inplace=True)
        y = F.relu(self.bn(self.conv_conv(y)), inplace=True)
        y = self.bn_expand(self.conv_expand(y))
        return F.relu(self.shortcut(x) + y, inplace=True)


class CifarResNeXt(nn.Module):
    """resnext.py:67-142: (B, in_channels, 32, 32) -> (B, nlabels)."""

    def __init__(self, nlabels, cardinality=8, depth=29, base_width=64, widen_factor=4, in_channels=3):
        super().__init__()
        self.cardinality, self.depth, self.base_width, self.widen_factor = cardinality, depth, base_width, widen_factor
        self.block_depth = (depth - 2) // 9
        self.nlabels = nlabels
        self.stages = [64, 64 * widen_factor, 128 * widen_factor, 256 * widen_factor]
        self.conv_1_3x3 = nn.Conv2d(in_channels, 64, 3, 1, 1, bias=False)
        self.bn_1 = nn.BatchNorm2d(64)
        self.stage_1 = self._stage("stage_1", self.stages[0], self.stages[1], 1)
        self.stage_2 = self._stage("stage_2", self.stages[1], self.stages[2], 2)
        self.stage_3 = self._stage("stage_3", self.stages[2], self.stages[3], 2)
        self.classifier = nn.Linear(self.stages[3], nlabels)
        init.kaiming_normal_(self.classifier.weight)
        for key, value in self.state_dict().items():
            leaf = key.split(".")[-1]
            if leaf == "weight":
                if "conv" in key:
                    init.kaiming_normal_(value, mode="fan_out")
                if "bn" in key:
                    value[...] = 1
            elif leaf == "bias":
                value[...] = 0

    def _stage(self, name, cin, cout, stride):
        block = nn.Sequential()
        for j in range(self.block_depth):
            block.add_module("%s_bottleneck_%d" % (name, j),
                             ResNeXtBottleneck(cin if j == 0 else cout, cout, stride if j == 0 else 1,
                                               self.cardinality, self.base_width, self.widen_factor))
        return block

    def forward(self, x):
        x = F.relu(self.bn_1(self.conv_1_3x3(x)), inplace=True)
        x = self.stage_3(self.stage_2(self.stage_1(x)))
        x = F.avg_pool2d(x, 8, 1)
        return self.classifier(x.view(-1, self.stages[3]))


def _bias_act(y, bias_f32, res=None, relu=True):
    """y <- relu?(y + bias[c] (+ res)) in place.  bf16 channels-last CUDA activations take ONE pass through
    ``ap_bias_act_nhwc_bf16`` (the bias add, residual add and clamp were 68 separate elementwise launches per batch);
    anything else (fp32 / CPU, used by the tests of the folding) takes the torch ops."""
    if (y.is_cuda and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
            and (res is None or (res.dtype == y.dtype and res.is_contiguous(memory_format=torch.channels_last)))):
        from . import _lib
        n, c, h, w = y.shape
        with torch.cuda.device(y.device):
            _lib.check(_lib.load().ap_bias_act_nhwc_bf16(y.data_ptr(), bias_f32.data_ptr(),
                                                         res.data_ptr() if res is not None else None,
                                                         n * h * w, c, 1 if relu else 0, _lib.stream_ptr()))
        return y
    y = y + bias_f32.to(y.dtype).reshape(1, -1, 1, 1)
    if res is not None:
        y = y + res
    return F.relu(y, inplace=True) if relu else y


def _conv_nobias(conv: nn.Conv2d, x):
    return F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)


class _FusedBottleneck(nn.Module):
    def __init__(self, src: ResNeXtBottleneck):
        super().__init__()
        self.reduce = _fold(src.conv_reduce, src.bn_reduce)
        self.conv = _fold(src.conv_conv, src.bn)
        self.expand = _fold(src.conv_expand, src.bn_expand)
        self.shortcut = None
        tail = self.expand.bias.detach().clone()
        if len(src.shortcut) > 0:
            self.shortcut = _fold(src.shortcut.shortcut_conv, src.shortcut.shortcut_bn)
            tail = tail + self.shortcut.bias.detach()  # the two biases meet in the residual add: applied once
        # fp32 bias tables for the fused epilogue (buffers: they stay fp32 when the convs are cast to bf16)
        self.register_buffer("b_reduce", self.reduce.bias.detach().float().clone())
        self.register_buffer("b_conv", self.conv.bias.detach().float().clone())
        self.register_buffer("b_tail", tail.float())

    def forward(self, x):
        y = _bias_act(_conv_nobias(self.reduce, x), self.b_reduce)
        y = _bias_act(_conv_nobias(self.conv, y), self.b_conv)
        r = x if self.shortcut is None else _conv_nobias(self.shortcut, x)
        return _bias_act(_conv_nobias(self.expand, y), self.b_tail, res=r)


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    """conv (no bias) + eval-mode batch-norm -> one conv with bias (exact in real arithmetic)."""
    out = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                    groups=conv.groups, bias=True)
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    out.weight.data.copy_(conv.weight.detach() * scale.reshape(-1, 1, 1, 1))
    out.bias.data.copy_(bn.bias.detach() - bn.running_mean.detach() * scale)
    return out


class FusedResNeXt(nn.Module):
    """Inference form of a trained/loaded ``CifarResNeXt``: batch-norms folded into the
    convolutions, channels-last, bf16 weights and activations with fp32 accumulation (cuDNN), fp32 logits.
    ``FusedResNeXt(clf)`` stands in for ``clf.eval()`` in the ``classifier`` slot for INFERENCE ONLY (evaluation,
    certification, black-box queries): it has no backward pass and raises if its input requires grad -- keep the
    plain module for attacks that back-propagate through ``AcousticSystem``.  It removes the 31 batch-norm
    and ~50 layout-conversion launches per batch that the reference module issues, and applies bias, residual add
    and ReLU in one in-place pass per convolution (``ap_bias_act_nhwc_bf16``)."""

    def __init__(self, src: CifarResNeXt, dtype=torch.bfloat16):
        super().__init__()
        assert not src.training, "fold batch-norm statistics of an eval() module"
        self.dtype = dtype
        self.stem = _fold(src.conv_1_3x3, src.bn_1)
        self.blocks = nn.Sequential(*[_FusedBottleneck(b) for stage in (src.stage_1, src.stage_2, src.stage_3) for b in stage])
        self.classifier = src.classifier
        self.width = src.stages[3]
        self.register_buffer("b_stem", self.stem.bias.detach().float().clone())
        self.to(memory_format=torch.channels_last)
        for conv in [m for m in self.modules() if isinstance(m, nn.Conv2d)]:
            conv.to(dtype)  # the fp32 bias buffers of the fused epilogue are left alone

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError(
                "FusedResNeXt is inference-only (in-place bf16 epilogue kernels, no autograd): an input that requires "
                "grad would silently lose its gradient here.  Use the plain CifarResNeXt module in the classifier slot "
                "when back-propagating through AcousticSystem (adaptive attacks).")
        with torch.no_grad():
            return self._forward(x)

    def _forward(self, x):
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = _bias_act(_conv_nobias(self.stem, x), self.b_stem)
        x = self.blocks(x)
        x = F.avg_pool2d(x, 8, 1).float()
        return self.classifier(x.reshape(-1, self.width))
