"""Oracle (test infrastructure): randomized-smoothing certification, restated.

Restates ``robustness_eval/certified_robust.py`` with the noise injected:
wherever the reference draws ``torch.normal(0, sigma, ...)`` on the CPU
(:47) the caller passes standard-normal draws ``z`` and the oracle forms
``delta = sigma * z``.  Vote counting is integer arithmetic and is bit-exact.
"""

import math

import torch
from scipy.stats import beta as _beta
from scipy.stats import norm as _norm

from .purify import one_shot_denoise
from .schedule import compute_t_star


def vote_counts(logits, num_classes=None):
    """certified_robust.py:58-67: argmax per row, histogram into int64[num_classes]."""
    if num_classes is None:
        num_classes = logits.shape[-1]
    predictions = logits.max(1, keepdim=True)[1].squeeze(1)
    counts = torch.zeros(num_classes, dtype=torch.int64)
    for i in range(num_classes):
        counts[i] = (predictions == i).sum().item()
    return counts


def smooth_logits(hp, eps_fn, transform, classifier, x, z, sigma):
    """certified_robust.py:44-56 + 17-31 for one batch of draws.

    x: (1,L) clip, z: (b,1,L) standard normal.  Returns the (b,K) logits of
    classifier(transform(one_shot_denoise(sqrt(abar*) * (x + sigma z)))).
    """
    b = z.shape[0]
    x_in = x.repeat(b, 1, 1) + sigma * z
    if eps_fn is not None:
        alpha_bar_star, t_star = compute_t_star(hp["Alpha_bar"], sigma)
        x_in = alpha_bar_star ** 0.5 * x_in
        x_in = one_shot_denoise(hp, eps_fn, x_in, t_star)
    if transform is not None:
        x_in = transform(x_in)
    return classifier(x_in)


def smooth_predict(hp, eps_fn, transform, classifier, x, z, sigma, batch_size, num_classes=10):
    """certified_robust.py:33-67.  z: (num_sampling,1,L); batches of ``batch_size`` plus the remainder (:38-40)."""
    n = z.shape[0]
    outs = []
    for s in range(0, n, batch_size):
        outs.append(smooth_logits(hp, eps_fn, transform, classifier, x, z[s:s + batch_size], sigma))
    return vote_counts(torch.cat(outs, dim=0), num_classes)


def lower_conf_bound(k, n, alpha=0.001):
    """certified_robust.py:113-117: ``proportion_confint(k, n, alpha=2*alpha, method='beta')[0]``.

    statsmodels 0.13.2 (requirements.txt:10) is absent; its 'beta' method is the
    Clopper-Pearson interval, whose lower end is beta.ppf(alpha_ci/2, k, n-k+1)
    (0 when k == 0) with alpha_ci = 2*alpha.  Parity unpinned against statsmodels.
    """
    k = int(k)
    if k == 0:
        return 0.0
    return float(_beta.ppf(alpha, k, n - k + 1))


def certify_from_counts(counts_0, counts, n, sigma, alpha=0.001):
    """certified_robust.py:84-96: (class, radius) from the selection and estimation counts."""
    c_A = int(counts_0.max(0, keepdim=True)[1].item())
    pa = lower_conf_bound(int(counts[c_A]), n, alpha)
    if pa > 0.5:
        return c_A, float(sigma * _norm.ppf(pa))
    return -1, 0.0
