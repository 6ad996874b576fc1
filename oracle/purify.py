"""Oracle (test infrastructure): the purifiers, restated with injected noise.

Restates ``diffusion_models/diffwave_ddpm.py`` (DDPM purifier, one-shot
denoise) and ``diffusion_models/diffwave_sde.py`` (reverse VP-SDE purifier)
in fp32 CPU arithmetic.  Wherever the reference draws noise
(``torch.normal(...).cuda()`` at diffwave_ddpm.py:66,100; ``randn_like`` at
diffwave_sde.py:185; torchsde's Brownian increments) these functions take the
already-drawn standard normal tensors instead, in the reference's draw order,
so the CUDA path can be fed the identical values.

``eps_fn(x, t)`` is any callable returning eps_theta(x, t); the default is
``oracle.wavenet.eps_theta`` bound to a state dict.
"""

import math

import torch


def ddpm_diffuse(hp, x0, t_star, z):
    """diffwave_ddpm.py:49-73 (``_diffusion``): x_t = sqrt(abar)*x0 + sqrt(1-abar)*z at index t*-1."""
    abar = hp["Alpha_bar"][t_star - 1]
    return torch.sqrt(abar) * x0 + torch.sqrt(1 - abar) * z


def ddpm_coefficients(hp, eps_fn, x_t, t):
    """diffwave_ddpm.py:143-164 (``compute_coefficients``) -> (eps, mu, sigma)."""
    Alpha, Alpha_bar, Sigma = hp["Alpha"], hp["Alpha_bar"], hp["Sigma"]
    eps = eps_fn(x_t, t)
    mu = (x_t - (1 - Alpha[t]) / torch.sqrt(1 - Alpha_bar[t]) * eps) / torch.sqrt(Alpha[t])
    return eps, mu, Sigma[t]


def ddpm_reverse(hp, eps_fn, x_t, t_star, z_steps):
    """diffwave_ddpm.py:75-104 (``_reverse``).  ``z_steps[i]`` is the noise of the
    i-th loop iteration (t = t*-1-i); the last iteration (t == 0) draws none."""
    x = x_t.clone()
    i = 0
    for t in range(t_star - 1, -1, -1):
        _, mu, sigma = ddpm_coefficients(hp, eps_fn, x, t)
        if t > 0:
            x = mu + sigma * z_steps[i]
            i += 1
        else:
            x = mu
    return x


def ddpm_purify(hp, eps_fn, x0, t_star, z):
    """diffwave_ddpm.py:36-47 (``DiffWave.forward``).  ``z`` has shape (t*, B, 1, L):
    z[0] feeds ``_diffusion``, z[1:] the reverse steps, in the reference's draw order."""
    x_t = ddpm_diffuse(hp, x0, t_star, z[0])
    return ddpm_reverse(hp, eps_fn, x_t, t_star, z[1:])


def one_shot_denoise(hp, eps_fn, x_t, reverse_timestep):
    """diffwave_ddpm.py:174-182,195-205: x0_hat = sqrt(1/abar_t)*x - sqrt(1/abar_t - 1)*eps, t = reverse_timestep-1."""
    t = reverse_timestep - 1
    eps = eps_fn(x_t, t)
    Alpha_bar = hp["Alpha_bar"]
    sqrt_recip = (1 / Alpha_bar).sqrt()
    sqrt_recipm1 = (1 / Alpha_bar - 1).sqrt()
    return sqrt_recip[t].float() * x_t - sqrt_recipm1[t].float() * eps


# --- reverse VP-SDE -----------------------------------------------------------


def _scale_timesteps(tab, t):
    """diffwave_sde.py:69-71."""
    return (t.float() * tab["N"]).long()


def sde_f(tab, eps_fn, t, x, audio_shape=(1, 16000)):
    """diffwave_sde.py:118-125 (``RevVPSDE.f``) with 73-105 inlined; x is (B, prod(audio_shape)),
    t a 0-dim fp32 tensor of *solver* time (the reference flips it: t' = 1 - t)."""
    tt = (1 - t).expand(x.shape[0])
    disc = _scale_timesteps(tab, tt) - 1
    beta_t = tab["discrete_betas"][disc] * tab["N"]
    drift = -0.5 * beta_t[:, None] * x
    diffusion = torch.sqrt(beta_t)
    x_audio = x.view(-1, *audio_shape)
    eps = eps_fn(x_audio, disc[0]).view(x.shape[0], -1)
    score = -eps / tab["sqrt_1m_alphas_cumprod"][disc[0]]
    drift = drift - diffusion[:, None] ** 2 * score
    return -drift


def sde_g(tab, t, x):
    """diffwave_sde.py:127-134 (``RevVPSDE.g``) with 107-116 inlined."""
    tt = (1 - t).expand(x.shape[0])
    disc = _scale_timesteps(tab, tt) - 1
    beta_t = tab["discrete_betas"][disc] * tab["N"]
    diffusion = torch.sqrt(beta_t)
    if disc.unique() > 0:
        ac = tab["alphas_cumprod"]
        scale = torch.sqrt(1 - ac[disc - 1]) / torch.sqrt(1 - ac[disc])
    else:
        scale = 0
    diffusion = scale * diffusion
    return diffusion[:, None].expand(x.shape)


def sde_purify(tab, eps_fn, x0, t, e, z_steps, T=200):
    """diffwave_sde.py:167-212 for ``sample_step == 1``, ``rand_t == False``.

    Diffuse with ``e`` (the ``randn_like`` at :185) using the on-the-fly cumprod of
    :190-191, then integrate from t0 = 1 - t/T - 1e-5 with fixed dt = 1/T
    (:194-204).  torchsde 0.2.5 is absent, so its ``method='euler'`` is restated
    from its published algorithm (Euler-Maruyama, y <- y + f(t,y) dt + g(t,y) dW
    with dW = sqrt(dt) * z): exactly ``t`` steps at indices k = t-1 .. 0.
    ``z_steps`` has shape (t, B, L); z_steps[i] multiplies the i-th increment.
    """
    B = x0.shape[0]
    betas = tab["discrete_betas"].float()
    a = (1 - betas).cumprod(dim=0)
    x = x0 * a[t - 1].sqrt() + e * (1.0 - a[t - 1]).sqrt()
    t0 = 1 - t / T + (-1e-5)
    dt = 1.0 / T
    sqrt_dt = math.sqrt(dt)
    y = x.view(B, -1)
    audio_shape = tuple(x0.shape[1:])
    for i in range(t):
        tc = torch.tensor(t0 + i * dt, dtype=torch.float32)
        f = sde_f(tab, eps_fn, tc, y, audio_shape)
        g = sde_g(tab, tc, y)
        y = y + f * dt + g * (z_steps[i] * sqrt_dt)
    return y.view(x0.shape)
