"""Oracle (test infrastructure): EOT and NES query loops of the black-box attacks, restated with injected noise.

Restates ``robustness_eval/_EOT.py:17-69`` (forward-only: ``use_grad=False``, as the black-box callers build it,
``black_box_attack.py:180-184``) and ``robustness_eval/_NES.py:15-55``.  Where the reference draws
``torch.randn([n_audios, S // 2, n_channels, N])`` (_NES.py:19) the caller passes the standard-normal tensor.
Pinned against the unmodified reference classes by ``tests/golden/nes.npz`` (oracle/make_golden.py ``nes``).
"""

from collections import Counter

import numpy as np
import torch


def eot(model, loss_fn, x_batch, y_batch, EOT_size, EOT_batch_size):
    """_EOT.py:17-69 with use_grad=False -> (scores, loss, decisions)."""
    num_batches = EOT_size // EOT_batch_size
    n_audios = x_batch.shape[0]
    scores = loss = None
    decisions = [[] for _ in range(n_audios)]
    for _ in range(num_batches):
        xr = x_batch.repeat(EOT_batch_size, 1, 1)
        yr = y_batch.repeat(EOT_batch_size)
        s = model(xr)
        d = s.max(1, keepdim=True)[1]
        l = loss_fn(s, yr)
        s_mean = s.view(EOT_batch_size, -1, s.shape[1]).mean(0)
        l_mean = l.view(EOT_batch_size, -1).mean(0)
        scores = s_mean if scores is None else scores + s_mean
        loss = l_mean if loss is None else loss + l_mean
        d = d.view(EOT_batch_size, -1).numpy()
        for ii in range(n_audios):
            decisions[ii] += list(d[:, ii])
    return scores / num_batches, loss / num_batches, decisions


def resolve_prediction(decisions):
    """_utils.py:127-135."""
    return np.array([Counter(d).most_common(1)[0][0] for d in decisions])


def nes(model, loss_fn, x, y, z, samples_per_draw, samples_per_draw_batch, sigma, EOT_size=1, EOT_batch_size=1):
    """_NES.py:15-55.  z: (num_batches, n_audios, samples_per_draw_batch // 2, n_channels, N) standard normal.
    -> (mean_loss, grad, adver_loss, adver_score, predict)."""
    n_audios, n_channels, N = x.shape
    S = samples_per_draw_batch
    num_batches = samples_per_draw // S
    EOT_num_batches = EOT_size // EOT_batch_size
    for i in range(num_batches):
        noise = torch.cat((z[i], -z[i]), 1)                                   # :19-21
        if i == 0:
            noise = torch.cat((torch.zeros_like(x).unsqueeze(1), noise), 1)   # :22-23
        eval_input = (noise * sigma + x.unsqueeze(1)).view(-1, n_channels, N)  # :24-25
        eval_y = torch.cat([torch.full((S + 1 if i == 0 else S,), int(y_), dtype=torch.long) for y_ in y])  # :26-32
        scores, loss, decisions = eot(model, loss_fn, eval_input, eval_y, EOT_size, EOT_batch_size)
        loss = loss / EOT_num_batches                                         # :35 (a second division, kept)
        scores = scores / EOT_num_batches                                     # :36
        loss = loss.view(n_audios, -1)
        scores = scores.view(n_audios, -1, scores.shape[1])
        if i == 0:
            adver_loss = loss[..., 0]
            loss = loss[..., 1:]
            adver_score = scores[:, 0, :]
            noise = noise[:, 1:, :, :]
            grad = torch.mean(loss.unsqueeze(2).unsqueeze(3) * noise, 1)
            mean_loss = loss.mean(1)
            predict = resolve_prediction(decisions).reshape(n_audios, -1)[:, 0]
        else:
            grad = grad + torch.mean(loss.unsqueeze(2).unsqueeze(3) * noise, 1)
            mean_loss = mean_loss + loss.mean(1)
    grad = grad / sigma / num_batches
    mean_loss = mean_loss / num_batches
    return mean_loss, grad, adver_loss, adver_score, predict


def toy_model(n_in, n_classes=10, seed=77):
    """A small deterministic stand-in for the AcousticSystem in the NES/EOT fixtures: (B,1,N) -> (B,n_classes),
    tanh of a fixed random projection (numpy PCG64), so the reference classes can be run in milliseconds."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = torch.from_numpy(rng.normal(0.0, 1.0 / np.sqrt(n_in), size=(n_classes, n_in)).astype(np.float32))
    b = torch.from_numpy(rng.normal(0.0, 0.1, size=(n_classes,)).astype(np.float32))

    def f(x):
        return 3.0 * torch.tanh(x.reshape(x.shape[0], -1) @ w.to(x.device).t() + b.to(x.device))

    return f
