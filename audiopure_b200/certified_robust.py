"""Randomized-smoothing certification: drop-in for ``robustness_eval/certified_robust.py`` with the draw
loop on the GPU and, optionally, sharded over the GPUs of one box.

Same surface as the reference (``RobustCertificate(classifier, transform, denoiser, one_shot_rev,
num_classes)``, ``certify``, ``smooth_predict``, ``compute_t_star``, ``lower_conf_bound``).  What changes:

* the smoothing noise is drawn on the device by a counter-based Philox keyed on (seed, clip, draw index,
  sample), fused with the ``x + delta`` and ``sqrt(alpha_bar*)`` scaling (certified_robust.py:46-54) -- the
  reference draws on the CPU and copies 64 KB per draw over PCIe; pre-drawn noise can be injected (``z=``);
* votes are counted on the device into int64 counters (certified_robust.py:58-67) and read back once;
* with ``world_size > 1`` each rank takes a contiguous slice of the draw indices and only the vote counts
  are all-reduced (NCCL via the C ABI, or any ``allreduce`` callable -- gloo in the CPU tests).  Because the
  noise is keyed on the draw index, the draws -- and therefore the summed integer counts -- do not depend
  on the number of ranks.
"""

import ctypes

import torch
from scipy.stats import beta as _beta_dist
from scipy.stats import norm

from . import _lib


def shard_range(n, rank, world_size):
    """Contiguous slice [lo, hi) of n draw indices for ``rank``; the remainder goes to the low ranks."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class NcclCountsAllReduce:
    """Sums int64 device counters over ranks with ``ap_allreduce_counts``.  The NCCL unique id is created on
    rank 0 and shipped through an already-initialised ``torch.distributed`` group (any backend)."""

    def __init__(self, rank, world_size):
        import torch.distributed as dist

        self.lib = _lib.load()
        buf = ctypes.create_string_buffer(_lib.AP_COMM_ID_BYTES)
        if rank == 0:
            _lib.check(self.lib.ap_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0)
        self.comm = ctypes.c_void_p()
        _lib.check(self.lib.ap_comm_init(rank, world_size, box[0], ctypes.byref(self.comm)))

    def __call__(self, counts):
        assert counts.is_cuda and counts.dtype == torch.int64 and counts.is_contiguous()
        _lib.check(self.lib.ap_allreduce_counts(self.comm, counts.data_ptr(), counts.numel(), _lib.stream_ptr()))
        return counts

    def __del__(self):
        try:
            if getattr(self, "comm", None):
                self.lib.ap_comm_destroy(self.comm)
        except Exception:
            pass


def torch_counts_allreduce(counts):
    """All-reduce through torch.distributed (gloo on CPU tensors, nccl on CUDA tensors)."""
    import torch.distributed as dist

    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


class RobustCertificate():

    def __init__(self, classifier: torch.nn.Module, transform=None, denoiser=None, one_shot_rev: bool = False,
                 num_classes=10, seed: int = 0, rank: int = 0, world_size: int = 1, allreduce=None):
        self.classifier = classifier
        self.transform = transform
        self.denoiser = denoiser
        self.num_classes = num_classes
        self.one_shot_rev = one_shot_rev
        self.seed = seed
        self.rank, self.world_size = rank, world_size
        self.allreduce = allreduce
        if world_size > 1 and allreduce is None:
            raise ValueError("world_size > 1 needs an allreduce callable (NcclCountsAllReduce or torch_counts_allreduce)")

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """certified_robust.py:17-31."""
        x_in = x
        if self.denoiser is not None:
            x_in = self.denoiser.one_shot_denoise(x_in)
        if self.transform is not None:
            x_in = self.transform(x_in)
        return self.classifier(x_in)

    @torch.no_grad()
    def smooth_predict(self, x: torch.Tensor, num_sampling: int = 100, sigma=0.25, batch_size=64, z: torch.Tensor = None,
                       clip: int = 0, first_draw: int = 0):
        """certified_robust.py:33-67 -> int64 counts[num_classes] (CPU tensor, like the reference).

        ``z``: optional injected standard-normal draws (num_sampling, 1, L); ``clip`` / ``first_draw`` key the
        Philox stream so that every (clip, draw) pair has its own noise whatever the sharding or batching."""
        assert (x.shape[0] == 1)
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.AudioPureError("smooth_predict runs on a CUDA device only (no CPU fallback)")
        x = x.to(torch.float32).contiguous()
        L = x.shape[-1]
        lo, hi = shard_range(num_sampling, self.rank, self.world_size)
        scale = 1.0
        if self.denoiser is not None:
            alpha_bar_star = 1 / (1 + sigma ** 2)
            t_star = self.compute_t_star(alpha_bar_star)
            self.denoiser.reverse_timestep = t_star
            scale = alpha_bar_star ** 0.5
        counts = torch.zeros(self.num_classes, dtype=torch.int64, device=x.device)
        if z is not None:
            z = z.to(device=x.device, dtype=torch.float32).contiguous()
            assert z.shape[0] == num_sampling and z.shape[-1] == L
        with torch.cuda.device(x.device):
            for s in range(lo, hi, batch_size):
                b = min(batch_size, hi - s)
                x_in = torch.empty(b, 1, L, dtype=torch.float32, device=x.device)
                zb = z[s:s + b] if z is not None else None
                _lib.check(lib.ap_smooth_inputs(x.data_ptr(), L, b, float(sigma), float(scale),
                                                zb.data_ptr() if zb is not None else None, self.seed, clip,
                                                first_draw + s, x_in.data_ptr(), _lib.stream_ptr()))
                logits = self.forward(x_in).to(torch.float32).contiguous()
                assert logits.shape[-1] == self.num_classes
                _lib.check(lib.ap_vote_counts(logits.data_ptr(), b, self.num_classes, counts.data_ptr(),
                                              _lib.stream_ptr()))
            if self.world_size > 1:
                self.allreduce(counts)
        return counts.cpu()

    @torch.no_grad()
    def certify(self, x: torch.Tensor, y: torch.Tensor, sigma: float = 0.25, n_0: int = 100, n: int = 100000,
                alpha: float = 0.001, batch_size: int = 64, clip_offset: int = 0):
        """certified_robust.py:69-100 -> (y_pred, radius)."""
        y_pred, radius = -torch.ones_like(y), torch.zeros_like(y, dtype=torch.float32)
        for i in range(x.shape[0]):
            x_in = x[i]
            counts_0 = self.smooth_predict(x_in, num_sampling=n_0, sigma=sigma, batch_size=batch_size,
                                           clip=clip_offset + i, first_draw=0)
            c_A = counts_0.max(0, keepdim=True)[1].item()
            counts = self.smooth_predict(x_in, num_sampling=n, sigma=sigma, batch_size=batch_size,
                                         clip=clip_offset + i, first_draw=n_0)
            pa = self.lower_conf_bound(k=counts[c_A], n=n, alpha=alpha)
            if pa > 0.5:
                y_pred[i] = c_A
                radius[i] = sigma * norm.ppf(pa)
            else:
                y_pred[i] = -1
                radius[i] = 0
        return y_pred, radius

    def compute_t_star(self, alpha_bar_star):
        """certified_robust.py:102-110."""
        Alpha_bar = self.denoiser.diffusion_hyperparams["Alpha_bar"]
        return torch.abs(Alpha_bar - alpha_bar_star).min(0, keepdim=True)[1].item() + 1

    def lower_conf_bound(self, k, n, alpha=0.001):
        """certified_robust.py:113-117: Clopper-Pearson lower bound at level alpha (statsmodels'
        ``proportion_confint(k, n, alpha=2*alpha, method='beta')[0]`` == beta.ppf(alpha, k, n-k+1), 0 for k == 0)."""
        k = int(k)
        if k == 0:
            return 0.0
        return float(_beta_dist.ppf(alpha, k, n - k + 1))

    def certified_robust_correct(self, y_pred, y_target, r_c, r: float = 1.):
        """certified_robust.py:119-127."""
        correct = 0
        for i in range(len(y_pred)):
            if y_pred[i] == y_target[i] and r_c[i] >= r:
                correct += 1
        return correct
