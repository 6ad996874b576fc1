"""Reverse VP-SDE purifier: drop-in for ``diffusion_models/diffwave_sde.py`` (``RevVPSDE``,
``RevDiffWave``).

The reference hands ``RevVPSDE`` to ``torchsde.sdeint_adjoint(method='euler', dt=1/T)``
(diffwave_sde.py:199-204).  With its discrete-parameter drift/diffusion (:73-116) one Euler-Maruyama
step at index k is the affine update

    x <- (1 + beta_k/2) x - beta_k / sqrt(1 - abar_k) * eps_theta(x, k) + sqrt(beta_k) sqrt((1-abar_{k-1})/(1-abar_k)) z

so the integrator is the same fused "network evaluation + update" launch sequence as the DDPM chain with
other coefficients (``ap_sde_purify``), k = t-1 .. 0, exactly t steps (torchsde's fp32 time stepping can
add a degenerate zero-length step at some t -- SURVEY.md section 7 -- which is not reproduced).
``RevVPSDE.f`` / ``.g`` are kept for callers that drive their own solver.
"""

import numpy as np
import torch

from .diffwave_ddpm import DiffWave, create_diffwave_model, default_seed
from .schedule import sde_tables


class RevVPSDE(torch.nn.Module):
    """diffwave_sde.py:34-134."""

    def __init__(self, model: DiffWave, score_type="guided_diffusion", beta_min=0.02, beta_max=4, N=200,
                 audio_shape=(1, 16000), model_kwargs=None):
        super().__init__()
        self.model = model
        self.score_type = score_type
        self.model_kwargs = model_kwargs
        self.audio_shape = audio_shape
        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self.N = N
        self.discrete_betas, self.alphas, self.alphas_cumprod = sde_tables(N, beta_min, beta_max)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.noise_type = "diagonal"
        self.sde_type = "ito"
        if score_type != "guided_diffusion":
            raise NotImplementedError(f"Unknown score type in RevVPSDE: {score_type}!")  # diffwave_sde.py:102

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f"t has to be in [0, 1], but get {t} with shape {t.shape}"
        return (t.float() * self.N).long()

    def f(self, t, x):
        """diffwave_sde.py:118-125: solver-time drift (the network evaluation runs in the CUDA kernels)."""
        t = t.expand(x.shape[0])
        disc = self._scale_timesteps(1 - t) - 1
        k = int(disc[0])
        assert x.ndim == 2 and np.prod(self.audio_shape) == x.shape[1], x.shape
        beta_t = float(self.discrete_betas[k]) * self.N
        eps = self.model.compute_eps_t(x.view(-1, *self.audio_shape), k).view(x.shape[0], -1)
        score = -eps / float(self.sqrt_1m_alphas_cumprod[k])
        drift = -0.5 * beta_t * x - beta_t * score
        return -drift

    def g(self, t, x):
        """diffwave_sde.py:127-134."""
        t = t.expand(x.shape[0])
        disc = self._scale_timesteps(1 - t) - 1
        k = int(disc[0])
        beta_t = float(self.discrete_betas[k]) * self.N
        if k > 0:
            ac = self.alphas_cumprod
            scale = float(torch.sqrt(1 - ac[k - 1]) / torch.sqrt(1 - ac[k]))
        else:
            scale = 0.0
        return torch.full_like(x, scale * beta_t ** 0.5)


class _PurifyWithLinearisedGrad(torch.autograd.Function):
    """Forward: the CUDA purifier.  Backward: the Jacobian the reference's ``sdeint_adjoint`` sees.

    ``RevVPSDE.f`` evaluates the network through ``DiffWave.compute_eps_t``, which is ``@torch.no_grad()``
    (diffwave_ddpm.py:166), so the only term of the drift that carries gradient is ``0.5 * beta_t * x``
    (diffwave_sde.py:79,104,118-125), the diffusion ``g`` does not depend on x, and the initial diffusion is the
    scaling ``sqrt(abar_{t-1})`` (:190-191).  The input-Jacobian of one Euler-Maruyama purification is therefore
    the scalar ``sqrt(abar_{t-1}) * prod_k (1 + beta_k / 2)`` times the identity;
    supplying it here also spares the adjoint pass its second set of t network evaluations."""

    @staticmethod
    def forward(ctx, x, engine, t, z, seed, clip_offset, scale):
        ctx.scale = scale
        return engine.sde_purify(x, t, z=z, seed=seed, clip_offset=clip_offset)

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out * ctx.scale, None, None, None, None, None, None


class RevDiffWave(torch.nn.Module):
    """diffwave_sde.py:138-218.  ``args`` carries the reference's attribute names: ``ddpm_path``, ``ddpm_config``,
    ``t``, ``sample_step``, ``rand_t``, ``t_delta``, ``use_bm``, ``score_type`` (+ optional ``precision``:
    "bf16" | "tf32").  A ready ``DiffWave`` may be
    passed as ``model=`` instead of a checkpoint path."""

    def __init__(self, args, device=None, model: DiffWave = None, seed: int = None):
        super().__init__()
        self.args = args
        if device is None:
            device = torch.device("cuda")
        self.device = torch.device(device)
        audio_shape = (1, 16000)
        if model is None:
            model = create_diffwave_model(model_path=args.ddpm_path, config_path=args.ddpm_config,
                                          reverse_timestep=args.t, device=self.device,
                                          precision=getattr(args, "precision", "bf16"))
        model.eval().to(self.device)
        self.T = model.diffusion_hyperparams["T"]
        self.model = model
        self.rev_vpsde = RevVPSDE(model=model, score_type=getattr(args, "score_type", "guided_diffusion"),
                                  beta_min=0.0001 * self.T, beta_max=0.02 * self.T, N=self.T,
                                  audio_shape=audio_shape, model_kwargs=None)
        self.betas = self.rev_vpsde.discrete_betas.float().to(self.device)
        self.seed = default_seed() if seed is None else seed
        self._calls = 0

    def audio_editing_sample(self, audio, z: torch.Tensor = None, clip_offset: int = 0):
        """diffwave_sde.py:167-212.  ``z``: optional (sample_step, t+1, B, 1, L) -- per repeat: z[0] the diffusion
        noise ``e`` (:185), z[1+i] the Brownian increment of Euler-Maruyama step i divided by sqrt(dt)."""
        assert isinstance(audio, torch.Tensor)
        assert audio.ndim == 3, audio.ndim
        x0 = audio.to(self.device)
        eng = self.model.model.engine()
        xs = []
        for it in range(self.args.sample_step):
            total_noise_levels = self.args.t
            if getattr(self.args, "rand_t", False):
                total_noise_levels = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
            self._calls += 1
            seed = (self.seed * 0x9E3779B97F4A7C15 + 0x5DE00000 + self._calls) & 0xFFFFFFFFFFFFFFFF
            zi = None if z is None else z[it]
            if torch.is_grad_enabled() and x0.requires_grad:  # adaptive attack (white_box_attack.py:437-439)
                x0 = _PurifyWithLinearisedGrad.apply(x0, eng, total_noise_levels, zi, seed, clip_offset,
                                                     self.input_jacobian(total_noise_levels))
            else:
                x0 = eng.sde_purify(x0, total_noise_levels, z=zi, seed=seed, clip_offset=clip_offset)
            xs.append(x0)
        return torch.cat(xs, dim=0)

    def input_jacobian(self, t):
        """d(purified)/d(input) as a scalar: sqrt(abar_{t-1}) * prod_{k<t} (1 + beta_k / 2)."""
        b = self.rev_vpsde.discrete_betas.double()
        ac = self.rev_vpsde.alphas_cumprod.double()
        return float(ac[t - 1].sqrt() * torch.prod(1.0 + 0.5 * b[:t]))

    def forward(self, x, z: torch.Tensor = None):
        return self.audio_editing_sample(x, z=z)
