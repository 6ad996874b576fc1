"""Oracle (test infrastructure): diffusion schedule and step embedding.

Restates ``diffusion_models/DiffWave_Unconditional/util.py:68-123`` and the
SDE-side tables of ``diffusion_models/diffwave_sde.py:54-61`` on the CPU in
fp32, keeping the reference's evaluation order so the tables are bit-equal.
"""

import numpy as np
import torch


def calc_diffusion_hyperparams(T, beta_0, beta_T):
    """util.py:96-123 -- Beta/Alpha/Alpha_bar/Sigma as CPU fp32 (T,) tensors.

    The sequential in-place loop (util.py:115-117) is reproduced verbatim in
    order: Alpha_bar[t] and Beta_tilde[t] are updated in the same iteration.
    """
    Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    Sigma = torch.sqrt(Beta_tilde)
    return {"T": T, "Beta": Beta, "Alpha": Alpha, "Alpha_bar": Alpha_bar, "Sigma": Sigma}


def calc_diffusion_step_embedding(diffusion_steps, dim_in):
    """util.py:68-93 -- (B,1) float steps -> (B,dim_in) [sin | cos] embedding."""
    assert dim_in % 2 == 0
    half = dim_in // 2
    _embed = np.log(10000) / (half - 1)
    _embed = torch.exp(torch.arange(half) * -_embed)
    _embed = diffusion_steps * _embed
    return torch.cat((torch.sin(_embed), torch.cos(_embed)), 1)


def sde_tables(T=200, beta_min=0.0001 * 200, beta_max=0.02 * 200):
    """diffwave_sde.py:54-61 (RevVPSDE.__init__) with the arguments RevDiffWave
    passes at diffwave_sde.py:157-159: discrete betas and the *cumprod* flavour
    of alpha-bar (differs from util.py's sequential product in the last ulp).
    """
    discrete_betas = torch.linspace(beta_min / T, beta_max / T, T)
    alphas = 1.0 - discrete_betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    return {
        "N": T,
        "discrete_betas": discrete_betas,
        "alphas": alphas,
        "alphas_cumprod": alphas_cumprod,
        "sqrt_alphas_cumprod": torch.sqrt(alphas_cumprod),
        "sqrt_1m_alphas_cumprod": torch.sqrt(1.0 - alphas_cumprod),
    }


def compute_t_star(Alpha_bar, sigma):
    """certified_robust.py:51-52,102-110 -- smoothing sigma -> (alpha_bar*, t*)."""
    alpha_bar_star = 1 / (1 + sigma ** 2)
    t_star = torch.abs(Alpha_bar - alpha_bar_star).min(0, keepdim=True)[1].item() + 1
    return alpha_bar_star, t_star
