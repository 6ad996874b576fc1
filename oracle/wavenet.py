"""Oracle (test infrastructure): the DiffWave epsilon-network, restated.

Plain fp32 CPU restatement of
``diffusion_models/DiffWave_Unconditional/WaveNet.py`` operating directly on
a reference-layout state dict (no ``nn.Module``), with every intermediate
exposed so the CUDA kernels can be checked layer by layer.
"""

import math

import torch
import torch.nn.functional as F

from .schedule import calc_diffusion_step_embedding
from .weights import DEFAULT_WAVENET_CONFIG


def fold_weight_norm(g, v):
    """WaveNet.py:28,67,72 (``nn.utils.weight_norm``, dim=0): w = g * v / ||v||,
    the norm taken over (Cin, k) for each output channel."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
    return v * (g / norm)


def _wn(sd, prefix):
    return fold_weight_norm(sd[prefix + ".weight_g"], sd[prefix + ".weight_v"]), sd[prefix + ".bias"]


def swish(x):
    """WaveNet.py:10-11."""
    return x * torch.sigmoid(x)


def step_embedding(sd, cfg, steps):
    """WaveNet.py:123-126: sinusoid (util.py:68-93) -> fc_t1 -> swish -> fc_t2 -> swish.
    ``steps`` is the (B,1) float tensor the purifier builds at diffwave_ddpm.py:157."""
    e = calc_diffusion_step_embedding(steps, cfg["diffusion_step_embed_dim_in"])
    e = swish(F.linear(e, sd["residual_layer.fc_t1.weight"], sd["residual_layer.fc_t1.bias"]))
    e = swish(F.linear(e, sd["residual_layer.fc_t2.weight"], sd["residual_layer.fc_t2.bias"]))
    return e


def residual_block(sd, n, dilation, x, emb):
    """WaveNet.py:75-97.  Returns (next h, skip, shifted input, gate output).

    The in-place ``h += part_t`` (WaveNet.py:77-84) aliases ``x``, so the
    residual term at :97 is the *shifted* input.
    """
    p = "residual_layer.residual_blocks.%d" % n
    part_t = F.linear(emb, sd[p + ".fc_t.weight"], sd[p + ".fc_t.bias"])
    xs = x + part_t[:, :, None]
    w, b = _wn(sd, p + ".dilated_conv_layer.conv")
    h = F.conv1d(xs, w, b, dilation=dilation, padding=dilation)
    C = x.shape[1]
    out = torch.tanh(h[:, :C, :]) * torch.sigmoid(h[:, C:, :])
    wr, br = _wn(sd, p + ".res_conv")
    ws, bs = _wn(sd, p + ".skip_conv")
    res = F.conv1d(out, wr, br)
    skip = F.conv1d(out, ws, bs)
    return (xs + res) * math.sqrt(0.5), skip, xs, out


def eps_theta(sd, x, t, cfg=None, return_intermediates=False):
    """WaveNet.py:164-172 with ``diffusion_steps = t * ones((B,1))``
    (diffwave_ddpm.py:157,169,177).  x: (B,1,L) fp32, t: int or 0-dim tensor."""
    cfg = dict(DEFAULT_WAVENET_CONFIG, **(cfg or {}))
    B = x.shape[0]
    steps = t * torch.ones((B, 1))
    w0, b0 = _wn(sd, "init_conv.0.conv")
    h = torch.maximum(F.conv1d(x, w0, b0), torch.zeros(()))  # WaveNet.py:13-19,147
    emb = step_embedding(sd, cfg, steps)
    skip_sum = 0
    inter = {"h": [], "skip": [], "xs": [], "gate": []}
    for n in range(cfg["num_res_layers"]):
        d = 2 ** (n % cfg["dilation_cycle"])
        h, skip, xs, gate = residual_block(sd, n, d, h, emb)
        skip_sum = skip_sum + skip
        if return_intermediates:
            inter["h"].append(h)
            inter["skip"].append(skip)
            inter["xs"].append(xs)
            inter["gate"].append(gate)
    y = skip_sum * math.sqrt(1.0 / cfg["num_res_layers"])  # WaveNet.py:135
    wf, bf = _wn(sd, "final_conv.0.conv")
    y = F.relu(F.conv1d(y, wf, bf))
    y = F.conv1d(y, sd["final_conv.2.conv.weight"], sd["final_conv.2.conv.bias"])
    if return_intermediates:
        inter["skip_sum"] = skip_sum
        return y, inter
    return y
