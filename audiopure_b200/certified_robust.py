"""Randomized-smoothing certification: drop-in for ``robustness_eval/certified_robust.py`` with the draw
loop on the GPU and, optionally, sharded over the GPUs of one box.

Same surface as the reference (``RobustCertificate(classifier, transform, denoiser, one_shot_rev,
num_classes)``, ``certify``, ``smooth_predict``, ``compute_t_star``, ``lower_conf_bound``).  What changes:

* the smoothing noise is drawn on the device by a counter-based Philox keyed on (seed, clip, draw index,
  sample), fused with the ``x + delta`` and ``sqrt(alpha_bar*)`` scaling (certified_robust.py:46-54) -- the
  reference draws on the CPU and copies 64 KB per draw over PCIe; pre-drawn noise can be injected (``z=``);
* a ``certify`` call turns ALL its clips and both of its passes (n_0 selection draws, n estimation draws,
  certified_robust.py:81-93) into one clip-major work list of (clip, draw) items and runs it in full batches that
  may span clips -- the reference (and r01 of this package) ran clip by clip, pass by pass, which at 8 GPUs left
  12-draw slivers of a 64-batch for the n_0 = 100 pass;
* votes are counted on the device into int64 counters per (pass, clip) (certified_robust.py:58-67) and read back
  once per call; the Clopper-Pearson bound and the radius stay on the host in float64 like the reference;
* with ``world_size > 1`` the list's BATCHES are dealt out contiguously over the ranks and only the vote counts are
  all-reduced, once per call (NCCL via the C ABI, or any ``allreduce`` callable -- gloo in the CPU tests).  The
  noise is keyed on (clip, draw) and a draw sits in the same batch at the same row at every world size, so the
  summed integer counts do not depend on the number of ranks (all ranks must use the same ``batch_size``).
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


def work_batches(n_clips, per_clip, rank, world_size, batch_size):
    """The launch plan of one certify call on one rank.  The clip-major work list of ``n_clips * per_clip``
    (clip, draw) items is cut into batches of ``batch_size`` items FIRST -- full batches that may span clips, plus
    one ragged batch at the very end -- and the BATCHES are then dealt out contiguously over the ranks
    (``shard_range``).  A given item therefore sits in the same batch, at the same row, next to the same neighbours
    whatever the number of ranks, so even a consumer whose kernels are only reproducible per batch position (cuDNN /
    cuBLASLt stream-K style reductions) votes identically at every world size -- provided every process runs the same
    kernels (cuDNN autotuning may pick differently per process).  Every rank must use the same
    ``batch_size``.  Yields (first flat item, rows)."""
    total = n_clips * per_clip
    n_batches = (total + batch_size - 1) // batch_size
    lo, hi = shard_range(n_batches, rank, world_size)
    for j in range(lo, hi):
        yield j * batch_size, min(batch_size, total - j * batch_size)


def flat_to_clip_draw(flat, per_clip, first_draw=0):
    """(clip, draw index) of work item ``flat`` -- the rule the kernels apply (csrc smooth_inputs_kernel)."""
    return flat // per_clip, first_draw + flat % per_clip


class NcclCountsAllReduce:
    """Sums int64 device counters over ranks with ``ap_allreduce_counts``.  The NCCL unique id is created on
    rank 0 and shipped through an already-initialised ``torch.distributed`` group (any backend); a 1-rank
    communicator needs no group."""

    def __init__(self, rank, world_size):
        import torch.distributed as dist

        self.lib = _lib.load()
        buf = ctypes.create_string_buffer(_lib.AP_COMM_ID_BYTES)
        if rank == 0:
            _lib.check(self.lib.ap_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        if world_size > 1:
            dist.broadcast_object_list(box, src=0)
        self.comm = ctypes.c_void_p()
        _lib.check(self.lib.ap_comm_init(rank, world_size, box[0], ctypes.byref(self.comm)))

    def __call__(self, counts):
        assert counts.is_cuda and counts.dtype == torch.int64 and counts.is_contiguous()
        _lib.check(self.lib.ap_allreduce_counts(self.comm, counts.data_ptr(), counts.numel(), _lib.stream_ptr()))
        return counts

    def close(self):
        """Destroy the communicator now (otherwise at garbage collection)."""
        if getattr(self, "comm", None):
            self.lib.ap_comm_destroy(self.comm)
            self.comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def torch_counts_allreduce(counts):
    """All-reduce through torch.distributed (gloo on CPU tensors, nccl on CUDA tensors)."""
    import torch.distributed as dist

    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


class RobustCertificate():
    """``robustness_eval/certified_robust.py:6-127`` with the draw loop on the GPU.

    Noise keys.  Every smoothing draw is Philox noise keyed on (``seed``, clip key, draw index).  Clip keys advance
    with every call (a per-instance counter, like ``DiffWave._calls``): ``certify`` takes one fresh key per clip of
    its batch, ``smooth_predict`` one per call, so successive dataloader batches -- and the selection (n_0) and
    estimation (n) passes of one clip, which use disjoint draw indices -- never reuse a draw.  ``clip_offset=`` /
    ``clip=`` / ``first_draw=`` override the keys (sharding tests, reproducing a certificate).  ``seed=None`` takes
    ``torch.initial_seed()``: reseed torch (or pass ``seed``) to get different draws in another run.  With
    ``world_size > 1`` every rank must be built with the same seed and make the same sequence of calls."""

    def __init__(self, classifier: torch.nn.Module, transform=None, denoiser=None, one_shot_rev: bool = False,
                 num_classes=10, seed: int = None, rank: int = 0, world_size: int = 1, allreduce=None,
                 pad_batches: bool = True):
        self.classifier = classifier
        self.pad_batches = pad_batches
        self.transform = transform
        self.denoiser = denoiser
        self.num_classes = num_classes
        self.one_shot_rev = one_shot_rev
        self.seed = (torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.rank, self.world_size = rank, world_size
        self.allreduce = allreduce
        self._next_clip_key = 0
        self._x_in = None
        self.last_counts = None  # (counts_0, counts) of the last certify call, int64 CPU tensors [clips][classes]
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

    def _take_clip_keys(self, n, override):
        if override is not None:
            return int(override)
        key = self._next_clip_key
        self._next_clip_key += n
        return key

    def _scale_for(self, sigma):
        """certified_robust.py:50-54: retarget the denoiser to t*(sigma) and return sqrt(alpha_bar*)."""
        if self.denoiser is None:
            return 1.0
        alpha_bar_star = 1 / (1 + sigma ** 2)
        self.denoiser.reverse_timestep = self.compute_t_star(alpha_bar_star)
        return alpha_bar_star ** 0.5

    def _count_votes(self, x, per_clip, n_split, first_draw, sigma, batch_size, z, clip_key0):
        """Votes of every (clip, draw) pair, draw in [0, per_clip), of the clips ``x`` (C, L): the clip-major work
        list is cut into FULL batches that may span clips and the batches are dealt out over the ranks
        (``work_batches``), so neither a small n_0 nor many ranks leaves the GPU with a sliver of a batch.  Returns
        int64 device counts [2][C][K] (draws < n_split | draws >= n_split), summed over ranks with ONE all-reduce."""
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.AudioPureError("smooth_predict / certify run on a CUDA device only (no CPU fallback)")
        C, L = x.shape
        K = self.num_classes
        scale = self._scale_for(sigma)
        counts = torch.zeros(2, C, K, dtype=torch.int64, device=x.device)
        if z is not None:
            z = z.to(device=x.device, dtype=torch.float32).contiguous()
            assert z.numel() == C * per_clip * L, "injected noise must be (clips, draws, 1, L)"
        if self._x_in is None or self._x_in.shape != (batch_size, 1, L) or self._x_in.device != x.device:
            self._x_in = torch.zeros(batch_size, 1, L, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for s, b in work_batches(C, per_clip, self.rank, self.world_size, batch_size):
                # A ragged last batch still goes through the denoiser / transform / classifier as a FULL batch (its
                # tail rows hold the previous batch's inputs and their votes are not counted): every launch has one
                # shape, so tensor maps, cuDNN plans and autotuning are reused, and -- because the consumer's cuDNN
                # kernels are only reproducible per batch shape -- a draw's vote does not depend on how many ranks
                # or what batch boundaries the work list was cut into.
                x_in = self._x_in if self.pad_batches else self._x_in[:b]
                _lib.check(lib.ap_smooth_inputs_batch(x.data_ptr(), L, b, s, per_clip, first_draw, float(sigma),
                                                      float(scale), z.data_ptr() if z is not None else None,
                                                      self.seed, clip_key0, x_in.data_ptr(), _lib.stream_ptr()))
                logits = self.forward(x_in).to(torch.float32).contiguous()
                assert logits.shape == (x_in.shape[0], K)
                _lib.check(lib.ap_vote_counts_batch(logits.data_ptr(), b, K, s, per_clip, n_split, C,
                                                    counts.data_ptr(), _lib.stream_ptr()))
            if self.allreduce is not None:
                self.allreduce(counts)
        return counts

    @torch.no_grad()
    def smooth_predict(self, x: torch.Tensor, num_sampling: int = 100, sigma=0.25, batch_size=64, z: torch.Tensor = None,
                       clip: int = None, first_draw: int = 0):
        """certified_robust.py:33-67 -> int64 counts[num_classes] (CPU tensor, like the reference).

        ``z``: optional injected standard-normal draws (num_sampling, 1, L).  ``clip`` / ``first_draw`` override the
        Philox key of this call (default: a fresh clip key per call, draws 0..num_sampling-1)."""
        assert (x.shape[0] == 1)
        x = x.to(torch.float32).reshape(1, -1).contiguous()
        counts = self._count_votes(x, num_sampling, num_sampling, first_draw, sigma, batch_size, z,
                                   self._take_clip_keys(1, clip))
        return counts[0, 0].cpu()

    @torch.no_grad()
    def certify(self, x: torch.Tensor, y: torch.Tensor, sigma: float = 0.25, n_0: int = 100, n: int = 100000,
                alpha: float = 0.001, batch_size: int = 64, clip_offset: int = None, z: torch.Tensor = None):
        """certified_robust.py:69-100 -> (y_pred, radius).

        All clips of the call go through ONE work list: draws [0, n_0) of a clip are its selection pass
        (:84-87), draws [n_0, n_0 + n) its estimation pass (:89-93); the votes are counted on the device, summed
        over ranks once and read back once, then the Clopper-Pearson bound and the radius (:94-100) are evaluated
        on the host in float64 like the reference.  ``z``: optional injected draws (clips, n_0 + n, 1, L)."""
        C = x.shape[0]
        xs = x.to(torch.float32).reshape(C, -1).contiguous()
        counts = self._count_votes(xs, n_0 + n, n_0, 0, sigma, batch_size, z, self._take_clip_keys(C, clip_offset)).cpu()
        counts_0, counts_n = counts[0], counts[1]
        self.last_counts = (counts_0, counts_n)
        y_pred, radius = -torch.ones_like(y), torch.zeros_like(y, dtype=torch.float32)
        for i in range(C):
            c_A = counts_0[i].max(0, keepdim=True)[1].item()
            pa = self.lower_conf_bound(k=counts_n[i, c_A], n=n, alpha=alpha)
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
