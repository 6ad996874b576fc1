"""Batched query evaluation for the black-box / EOT callers of the purification path:
drop-ins for ``robustness_eval/_EOT.py`` (class ``EOT``) and ``robustness_eval/_NES.py`` (class ``NES``).

Both wrap an ``AcousticSystem`` and spend all their time in its forward pass -- ``samples_per_draw`` (x ``EOT_size``)
purify -> log-mel -> classify queries per attack iteration (``black_box_attack.py:180-184``,
``adaptive_attack_eval.py:153-160``).  What changes here:

* NES draws its antithetic perturbations on the device from a counter-based Philox stream keyed on
  (seed, audio, draw index) and writes ``x + sigma * noise`` directly (``ap_nes_inputs``); the gradient estimate
  ``mean_j loss_j * noise_j / sigma`` is reduced on the device by a kernel that RE-GENERATES the noise from its keys
  (``ap_nes_grad``) instead of keeping ``samples x L`` floats in HBM and multiplying them back in (_NES.py:19-24,44-48);
* EOT evaluates its ``EOT_size`` repetitions as ONE batch (the purifier already chunks internally) instead of
  ``EOT_size // EOT_batch_size`` sequential passes, so small ``EOT_batch_size`` values no longer starve the GPU;
* with ``world_size > 1`` NES shards the draw batches over ranks; Philox keys are (audio, draw) so the draws do not
  depend on the number of ranks, and only the (n_audios, L) gradient and the (n_audios,) loss are all-reduced.

Return values, argument names and the reference's arithmetic quirks (the second division of loss and scores by the
number of EOT batches at _NES.py:35-36) are kept so that an attack driven by these classes takes the same steps.
"""

from collections import Counter

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class EOT(nn.Module):
    """_EOT.py:5-69: expectation over ``EOT_size`` stochastic forward passes of ``model`` per input.

    -> (scores, loss, grad, decisions): scores / loss / grad averaged over the repetitions (_EOT.py:49-66),
    ``decisions[i]`` the list of the ``EOT_size`` top-1 decisions of input i (_EOT.py:57-59)."""

    def __init__(self, model, loss, EOT_size=1, EOT_batch_size=1, use_grad=True, max_rows=1024):
        super().__init__()
        self.model = model
        self.loss = loss
        self.EOT_size = EOT_size
        self.EOT_batch_size = EOT_batch_size
        self.EOT_num_batches = self.EOT_size // self.EOT_batch_size
        self.use_grad = use_grad
        self.max_rows = max_rows  # rows per model call (bounds activation memory of the consumer classifier)

    def forward(self, x_batch, y_batch, EOT_size=None, EOT_batch_size=None, use_grad=None):
        EOT_size = EOT_size if EOT_size else self.EOT_size
        EOT_batch_size = EOT_batch_size if EOT_batch_size else self.EOT_batch_size
        reps = (EOT_size // EOT_batch_size) * EOT_batch_size  # the reference drops the remainder (_EOT.py:22,32-36)
        use_grad = use_grad if use_grad else self.use_grad
        n_audios, n_channels, max_len = x_batch.size()
        per_call = max(1, self.max_rows // n_audios)
        scores = loss = grad = None
        decisions = [[] for _ in range(n_audios)]
        for r0 in range(0, reps, per_call):
            r = min(per_call, reps - r0)
            xr = x_batch.repeat(r, 1, 1)
            yr = y_batch.repeat(r)
            if use_grad:
                xr = xr.detach().requires_grad_(True)
                s = self.model(xr)
            else:
                with torch.no_grad():
                    s = self.model(xr)
            l = self.loss(s, yr)
            if use_grad:
                l.backward(torch.ones_like(l))
                g = xr.grad.view(r, -1, n_channels, max_len).sum(0)
                grad = g if grad is None else grad + g
            s, l = s.detach(), l.detach()
            d = s.max(1, keepdim=True)[1].view(r, -1).cpu().numpy()
            for ii in range(n_audios):
                decisions[ii] += list(d[:, ii])
            s = s.view(r, -1, s.shape[1]).sum(0)
            l = l.view(r, -1).sum(0)
            scores = s if scores is None else scores + s
            loss = l if loss is None else loss + l
        scores = scores / reps
        loss = loss / reps
        if grad is not None:
            grad = grad / reps
        return scores, loss, grad, decisions


def resolve_prediction(decisions):
    """_utils.py:127-135: majority vote over each input's decisions."""
    return np.array([Counter(d).most_common(1)[0][0] for d in decisions])


class NES(nn.Module):
    """_NES.py:6-55: natural-evolution-strategies gradient estimate of ``EOT_wrapper``'s loss with antithetic
    Gaussian perturbations.  ``forward(x (n_audios, 1, N), y) -> (mean_loss, grad, adver_loss, adver_score, predict)``.

    ``z=`` injects the standard-normal halves, shape (num_batches, n_audios, samples_per_draw_batch // 2, 1, N), for
    validation against the reference; by default they are Philox draws keyed on (seed, audio, draw index)."""

    def __init__(self, samples_per_draw, samples_per_draw_batch, sigma, EOT_wrapper, seed: int = None, rank: int = 0,
                 world_size: int = 1, allreduce=None):
        super().__init__()
        self.samples_per_draw = samples_per_draw
        self.samples_per_draw_batch_size = samples_per_draw_batch
        self.sigma = sigma
        self.EOT_wrapper = EOT_wrapper
        self.seed = (torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.rank, self.world_size, self.allreduce = rank, world_size, allreduce
        self._calls = 0
        if world_size > 1 and allreduce is None:
            raise ValueError("world_size > 1 needs an allreduce callable (e.g. torch.distributed.all_reduce)")

    def forward(self, x, y, z: torch.Tensor = None):
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.AudioPureError("NES runs on a CUDA device only (no CPU fallback)")
        n_audios, n_channels, N = x.shape
        assert n_channels == 1
        S = self.samples_per_draw_batch_size
        num_batches = self.samples_per_draw // S
        x = x.to(torch.float32).contiguous()
        y = torch.as_tensor(y, dtype=torch.long, device=x.device)
        if z is not None:
            z = z.to(device=x.device, dtype=torch.float32).contiguous()
            assert tuple(z.shape) == (num_batches, n_audios, S // 2, 1, N), tuple(z.shape)
        self._calls += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + self._calls) & 0xFFFFFFFFFFFFFFFF  # fresh draws every call
        EOT_num_batches = int(self.EOT_wrapper.EOT_size // self.EOT_wrapper.EOT_batch_size)
        grad = torch.zeros(n_audios, n_channels, N, device=x.device)
        mean_loss = torch.zeros(n_audios, device=x.device)
        adver_loss = adver_score = predict = None
        stream = _lib.stream_ptr
        with torch.cuda.device(x.device):
            for i in range(num_batches):
                if i != 0 and i % self.world_size != self.rank:
                    continue  # draw batch 0 (it carries the clean query) runs on every rank, the rest are dealt out
                lead = 1 if i == 0 else 0
                eval_input = torch.empty(n_audios * (lead + S), n_channels, N, device=x.device)
                zi = z[i] if z is not None else None
                _lib.check(lib.ap_nes_inputs(x.data_ptr(), n_audios, N, S, lead, float(self.sigma),
                                             zi.data_ptr() if zi is not None else None, seed, 0, i * (S // 2),
                                             eval_input.data_ptr(), stream()))
                eval_y = y.repeat_interleave(lead + S)                                       # _NES.py:26-32
                scores, loss, _, decisions = self.EOT_wrapper(eval_input, eval_y)
                loss = (loss / EOT_num_batches).view(n_audios, -1).contiguous()              # _NES.py:35,38
                scores = (scores / EOT_num_batches).view(n_audios, -1, scores.shape[1])      # _NES.py:36,39
                if i == 0:
                    adver_loss = loss[..., 0].clone()                                         # _NES.py:42
                    adver_score = scores[:, 0, :].clone()                                     # _NES.py:44
                    predict = resolve_prediction(decisions).reshape(n_audios, -1)[:, 0]      # _NES.py:48-49
                if i == 0 and self.rank != 0:
                    continue  # every rank needs the clean query's outputs; only rank 0 adds batch 0 to the sums
                # _NES.py:46,51: grad += mean_j(loss_j * noise_j); the / sigma / num_batches of :53 folded in
                _lib.check(lib.ap_nes_grad(loss.data_ptr(), lead + S, lead, n_audios, N, S,
                                           1.0 / (S * float(self.sigma) * num_batches),
                                           zi.data_ptr() if zi is not None else None, seed, 0, i * (S // 2),
                                           grad.data_ptr(), stream()))
                mean_loss += loss[..., lead:].mean(1) / num_batches                            # _NES.py:47,52,54
            if self.world_size > 1:
                self.allreduce(grad)
                self.allreduce(mean_loss)
        return mean_loss, grad, adver_loss, adver_score, predict
