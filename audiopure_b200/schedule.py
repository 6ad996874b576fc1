"""Host-side precompute: diffusion schedule and step embedding.

Mirrors ``diffusion_models/DiffWave_Unconditional/util.py:68-123`` of the reference (same names,
same evaluation order, so the tables are bit-equal) and the SDE-side tables of
``diffusion_models/diffwave_sde.py:54-61``.  These run once per model on the CPU; the per-step work
is in the CUDA kernels.
"""

import numpy as np
import torch


def calc_diffusion_hyperparams(T, beta_0, beta_T):
    """util.py:96-123.  Returns {"T", "Beta", "Alpha", "Alpha_bar", "Sigma"} with CPU fp32 (T,) tensors."""
    Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    Sigma = torch.sqrt(Beta_tilde)
    return {"T": T, "Beta": Beta, "Alpha": Alpha, "Alpha_bar": Alpha_bar, "Sigma": Sigma}


def calc_diffusion_step_embedding(diffusion_steps, diffusion_step_embed_dim_in):
    """util.py:68-93 (without the hard-coded ``.cuda()``): (B,1) steps -> (B,dim) [sin | cos]."""
    assert diffusion_step_embed_dim_in % 2 == 0
    half_dim = diffusion_step_embed_dim_in // 2
    _embed = np.log(10000) / (half_dim - 1)
    _embed = torch.exp(torch.arange(half_dim) * -_embed).to(diffusion_steps.device)
    _embed = diffusion_steps * _embed
    return torch.cat((torch.sin(_embed), torch.cos(_embed)), 1)


def sde_tables(T=200, beta_min=0.0001 * 200, beta_max=0.02 * 200):
    """diffwave_sde.py:57-61 as RevDiffWave configures it (:157-159)."""
    discrete_betas = torch.linspace(beta_min / T, beta_max / T, T)
    alphas = 1.0 - discrete_betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    return discrete_betas, alphas, alphas_cumprod
