"""DDPM purifier: drop-in for ``diffusion_models/diffwave_ddpm.py`` (class ``DiffWave``, factory
``create_diffwave_model``) whose network evaluations, reverse-step updates and noise run in the sm_100a
kernels.  Same constructor, attributes (``model``, ``diffusion_hyperparams``, mutable
``reverse_timestep``, ``freeze``) and methods as the reference.

Noise: the reference draws ``torch.normal`` on the CPU and copies it over (diffwave_ddpm.py:66,100).
Here the default is in-kernel Philox keyed on (seed, step, clip, sample) -- a different stream, same
distribution -- and every method accepts the already-drawn noise (``z=``) for bit-for-bit comparable
validation against the reference.
"""

import json
from typing import Union

import numpy as np
import torch

from .schedule import calc_diffusion_hyperparams
from .wavenet import WaveNet_Speech_Commands


def default_seed():
    """Philox seed when the caller gives none: torch's global seed (so ``torch.manual_seed`` governs the purifier's
    noise as it governs the reference's ``torch.normal`` draws), mixed with the rank when ``torch.distributed`` is
    initialised so that ranks of one job do not purify with identical noise."""
    seed = torch.initial_seed()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        seed ^= (torch.distributed.get_rank() + 1) * 0x9E3779B97F4A7C15
    return seed & 0xFFFFFFFFFFFFFFFF


def _as_tensor(x):
    if isinstance(x, np.ndarray):  # diffwave_ddpm.py:38-39,52-53,78-79
        x = torch.from_numpy(x)
    return x


class DiffWave(torch.nn.Module):

    def __init__(self, model: WaveNet_Speech_Commands, diffusion_hyperparams: dict, reverse_timestep: int = 200,
                 grad_enable=True, seed: int = None):
        super().__init__()
        self.model = model
        self.diffusion_hyperparams = diffusion_hyperparams
        self.reverse_timestep = reverse_timestep
        self.freeze = False
        self.grad_enable = grad_enable
        self.seed = default_seed() if seed is None else seed
        self._calls = 0
        self._check_tables()

    def _check_tables(self):
        hp = self.diffusion_hyperparams
        T = hp["T"]
        assert len(hp["Alpha"]) == T and len(hp["Alpha_bar"]) == T and len(hp["Sigma"]) == T  # diffwave_ddpm.py:59-61
        own = calc_diffusion_hyperparams(**self.model.diffusion_config)
        for k in ("Alpha", "Alpha_bar", "Sigma"):
            if not torch.equal(torch.as_tensor(hp[k]).float().cpu(), own[k]):
                raise ValueError("diffusion_hyperparams[%r] differs from the schedule the network was packed with; "
                                 "pass diffusion_config= to WaveNet_Speech_Commands" % k)

    def _next_seed(self):
        # a fresh Philox key per forward call, deterministic given self.seed
        self._calls += 1
        return (self.seed * 0x9E3779B97F4A7C15 + self._calls) & 0xFFFFFFFFFFFFFFFF

    def forward(self, waveforms: Union[torch.Tensor, np.ndarray], z: torch.Tensor = None, clip_offset: int = 0):
        """diffwave_ddpm.py:36-47: diffuse to ``reverse_timestep`` and run the reverse chain.
        ``z``: optional (reverse_timestep, B, 1, L) noise in the reference's draw order."""
        waveforms = _as_tensor(waveforms)
        assert waveforms.ndim == 3
        with torch.no_grad():
            return self.model.engine().ddpm_purify(waveforms, self.reverse_timestep, z=z, seed=self._next_seed(),
                                                   clip_offset=clip_offset)

    def _diffusion(self, x_0, z: torch.Tensor = None):
        """diffwave_ddpm.py:49-73."""
        x_0 = _as_tensor(x_0)
        assert x_0.ndim == 3
        eng = self.model.engine()
        ab = float(self.diffusion_hyperparams["Alpha_bar"][self.reverse_timestep - 1])
        x_0 = x_0.to(eng.device, torch.float32)
        if z is None:
            z = torch.randn_like(x_0)
        return (ab ** 0.5) * x_0 + ((1 - ab) ** 0.5) * z.to(eng.device)

    def _reverse(self, x_t, z: torch.Tensor = None):
        """diffwave_ddpm.py:75-104; ``z``: optional (reverse_timestep-1, B, 1, L)."""
        x = _as_tensor(x_t)
        assert x.ndim == 3
        i = 0
        for t in range(self.reverse_timestep - 1, -1, -1):
            zi = None
            if t > 0 and z is not None:
                zi = z[i]
                i += 1
            x = self._reverse_step(x, t, zi)
        return x

    def _coefs(self, t):
        hp = self.diffusion_hyperparams
        al, ab = float(hp["Alpha"][t]), float(hp["Alpha_bar"][t])
        return 1.0 / al ** 0.5, -(1 - al) / (1 - ab) ** 0.5 / al ** 0.5, float(hp["Sigma"][t])

    def _reverse_step(self, x, t, z=None):
        ca, cb, sig = self._coefs(t)
        return self.model.engine().step(x, t, ca, cb, sig if t > 0 else 0.0, z=z, seed=self._next_seed(), stream_id=t)

    def compute_coefficients(self, x_t, t: int):
        """diffwave_ddpm.py:143-164 -> (eps_theta, mu_theta, sigma_theta)."""
        x_t = _as_tensor(x_t)
        eng = self.model.engine()
        eps = eng.eps(x_t, t)
        ca, cb, sig = self._coefs(t)
        mu = ca * x_t.to(eng.device, torch.float32) + cb * eps
        return eps, mu, self.diffusion_hyperparams["Sigma"][t]

    @torch.no_grad()
    def compute_eps_t(self, x_t, t):
        """diffwave_ddpm.py:166-172; ``t`` may be an int or a 0-dim tensor."""
        return self.model.engine().eps(_as_tensor(x_t), int(t))

    def one_shot_denoise(self, x_t):
        """diffwave_ddpm.py:174-182."""
        return self.model.engine().one_shot(_as_tensor(x_t), self.reverse_timestep)


def create_diffwave_model(model_path, config_path, reverse_timestep=25, device="cuda", precision="bf16"):
    """diffwave_ddpm.py:395-411.  ``model_path=None`` keeps the random init (no checkpoint offline).
    ``precision``: "bf16" (default) or "tf32" -- the tensor-core mode of the residual-stack GEMMs."""
    with open(config_path) as f:
        cfg = json.loads(f.read())
    wavenet_config = cfg["wavenet_config"]
    diffusion_config = cfg["diffusion_config"]
    diffusion_hyperparams = calc_diffusion_hyperparams(**diffusion_config)
    WaveNet_model = WaveNet_Speech_Commands(**wavenet_config, diffusion_config=diffusion_config,
                                            precision=precision).to(device)
    if model_path is not None:
        checkpoint = torch.load(model_path, map_location="cpu")
        WaveNet_model.load_state_dict(checkpoint["model_state_dict"])
    return DiffWave(model=WaveNet_model, diffusion_hyperparams=diffusion_hyperparams, reverse_timestep=reverse_timestep)
