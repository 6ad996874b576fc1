"""The DiffWave epsilon-network as a drop-in for the reference's ``WaveNet_Speech_Commands``
(``diffusion_models/DiffWave_Unconditional/WaveNet.py:138-172``).

The module owns its parameters in the reference's *state-dict layout* (408 tensors for the shipped
config: ``...conv.weight_g`` / ``...conv.weight_v`` / ``bias`` for the weight-normed convs), so a
reference checkpoint loads with ``load_state_dict`` (``diffwave_ddpm.py:406-407``).  It holds no compute:
``forward`` hands the waveform to the sm_100a kernels through the C ABI.  Before the first evaluation
(and after any ``load_state_dict``) the weights are packed once -- weight-norm folded, GEMM operands
re-laid out and cast to bf16 (or rounded to tf32 with ``precision="tf32"``), the step-embedding MLP and every layer's ``fc_t`` tabulated for all T
steps -- see ``pack_weights``.
"""

import ctypes
import math

import torch
import torch.nn as nn

from . import _lib
from .schedule import calc_diffusion_hyperparams, calc_diffusion_step_embedding, sde_tables

DEFAULT_DIFFUSION_CONFIG = {"T": 200, "beta_0": 0.0001, "beta_T": 0.02}
PRECISIONS = ("bf16", "tf32")


def round_to_tf32(x):
    """fp32 -> nearest tf32 (10 mantissa bits, ties away from zero), returned as fp32 words with the low 13 bits
    clear: what the tensor core sees of an operand in the tf32 build (see ``tile_store32`` in csrc)."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class _WNConv(nn.Module):
    """Parameter holder with the key names of ``nn.utils.weight_norm(nn.Conv1d)`` (WaveNet.py:23-34)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        conv = nn.Conv1d(cin, cout, k)  # torch's default init == what the reference leaves in weight_v
        v = conv.weight.detach()
        self.bias = nn.Parameter(conv.bias.detach().clone())
        self.weight_g = nn.Parameter(v.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1).clone())
        self.weight_v = nn.Parameter(v.clone())

    def folded(self):
        v = self.weight_v.detach().float()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        return v * (self.weight_g.detach().float() / norm), self.bias.detach().float()


class _Wrap(nn.Module):
    """``Conv`` / ``ZeroConv1d`` wrappers put the conv under ``.conv`` (WaveNet.py:27,42)."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class _PlainConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, 1))  # ZeroConv1d: zero init (WaveNet.py:43-44)
        self.bias = nn.Parameter(torch.zeros(cout))


class _Block(nn.Module):
    def __init__(self, C, S, emb_out):
        super().__init__()
        self.fc_t = nn.Linear(emb_out, C)
        self.dilated_conv_layer = _Wrap(_WNConv(C, 2 * C, 3))
        self.res_conv = _WNConv(C, C, 1)
        self.skip_conv = _WNConv(C, S, 1)


class _Group(nn.Module):
    def __init__(self, C, S, layers, emb_in, emb_mid, emb_out):
        super().__init__()
        self.fc_t1 = nn.Linear(emb_in, emb_mid)
        self.fc_t2 = nn.Linear(emb_mid, emb_out)
        self.residual_blocks = nn.ModuleList([_Block(C, S, emb_out) for _ in range(layers)])


def _swish(x):
    return x * torch.sigmoid(x)


class WaveNet_Speech_Commands(nn.Module):
    """Same constructor and ``forward((audio, diffusion_steps))`` contract as WaveNet.py:138-172."""

    def __init__(self, in_channels=1, res_channels=256, skip_channels=128, out_channels=1, num_res_layers=30,
                 dilation_cycle=10, diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512,
                 diffusion_step_embed_dim_out=512, diffusion_config=None, max_chunk=None, precision="bf16"):
        super().__init__()
        if not (in_channels == 1 and out_channels == 1 and res_channels == 256 and skip_channels == 256):
            raise NotImplementedError(
                "the sm_100a kernels are built for in/out channels 1 and res/skip channels 256 "
                "(configs/config.json); got in=%d res=%d skip=%d out=%d" %
                (in_channels, res_channels, skip_channels, out_channels))
        self.num_res_layers = num_res_layers
        self.dilation_cycle = dilation_cycle
        self.embed_dim_in = diffusion_step_embed_dim_in
        self.diffusion_config = dict(diffusion_config or DEFAULT_DIFFUSION_CONFIG)
        if precision not in PRECISIONS:
            raise NotImplementedError("precision must be one of %s, got %r" % (PRECISIONS, precision))
        self.precision = precision      # "bf16" (default) or "tf32": include/audiopure_b200.h AP_FLAG_TF32
        self.max_chunk = max_chunk or (32 if precision == "tf32" else 64)  # clips per pass (bounds the workspace)
        self.init_conv = nn.Sequential(_Wrap(_WNConv(in_channels, res_channels, 1)))
        self.residual_layer = _Group(res_channels, skip_channels, num_res_layers, diffusion_step_embed_dim_in,
                                     diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        final = nn.Sequential()
        final.add_module("0", _Wrap(_WNConv(skip_channels, skip_channels, 1)))
        final.add_module("2", _Wrap(_PlainConv(skip_channels, out_channels)))
        self.final_conv = final
        self._engine = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # ------------------------------------------------------------------ packing / engine lifetime --
    def invalidate(self):
        """Drop the packed weights (call after mutating parameters in place)."""
        self._engine = None

    def _signature(self):
        p = next(self.parameters())
        return p.device, p.dtype, p.data_ptr()

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float() re-home the parameters: the packed copies follow.  A no-op move (same device and
        # dtype, e.g. RevDiffWave's `model.eval().to(device)` on a model that is already there) keeps them.
        before = self._signature()
        out = super()._apply(fn, *a, **k)
        if self._signature() != before:
            self._engine = None
        return out

    def engine(self):
        if self._engine is None:
            self._engine = Engine(self)
        return self._engine

    @torch.no_grad()
    def pack_weights(self, device):
        """Fold, re-lay-out and tabulate everything the kernels need (include/audiopure_b200.h: ap_weights)."""
        L, T = self.num_res_layers, self.diffusion_config["T"]
        C = 256
        dev = torch.device(device)
        f32 = dict(dtype=torch.float32, device=dev)
        blocks = self.residual_layer.residual_blocks

        # gate-interleaved row order of the dilated conv: chunk c holds gate channels [128c, 128c+128):
        # first their tanh rows (conv out channels 128c..), then their sigmoid rows (256+128c..)  (WaveNet.py:90)
        perm = torch.cat([torch.arange(128) + 128 * c + 256 * half for c in range(2) for half in range(2)]).to(dev)

        # the sigmoid rows carry a factor 1/2 (exact in bf16 / tf32): sigmoid(s) = 1/2 tanh(s/2) + 1/2 then needs no
        # multiply in the epilogue (csrc/sm100.cuh gate_act)
        half = torch.ones(512, 1, **f32)
        half[128:256] = 0.5
        half[384:512] = 0.5
        w1 = torch.empty(L, 512, 768, **f32)
        b1 = torch.empty(L, 512, **f32)
        w2 = torch.empty(L, 256, 256, **f32)
        b_res = torch.empty(L, C, **f32)
        ws = torch.empty(256, L * 256, **f32)
        bs = torch.zeros(256, **f32)
        rs = math.sqrt(1.0 / L)  # WaveNet.py:135
        for n, blk in enumerate(blocks):
            w, b = blk.dilated_conv_layer.conv.folded()          # (512, 256, 3)
            w = w.to(dev).permute(0, 2, 1).reshape(512, 768)     # K index = tap*256 + cin
            w1[n] = w[perm] * half
            b1[n] = b.to(dev)[perm] * half[:, 0]
            wr, br = blk.res_conv.folded()
            w2[n] = wr.to(dev)[:, :, 0] * math.sqrt(0.5) * 0.5   # WaveNet.py:97; x 1/2: the kernels keep 2 x gate
            b_res[n] = br.to(dev)
            wsk, bsk = blk.skip_conv.folded()
            ws[:, n * 256:(n + 1) * 256] = wsk.to(dev)[:, :, 0] * (rs * 0.5)   # x 1/2: see above
            bs += bsk.to(dev) * rs

        # step embedding for every t (util.py:68-93, WaveNet.py:124-126) and every layer's fc_t (WaveNet.py:82)
        steps = torch.arange(T, **f32).reshape(T, 1)
        g = self.residual_layer
        emb = calc_diffusion_step_embedding(steps, self.embed_dim_in)
        emb = _swish(torch.nn.functional.linear(emb, g.fc_t1.weight.detach().float().to(dev), g.fc_t1.bias.detach().float().to(dev)))
        emb = _swish(torch.nn.functional.linear(emb, g.fc_t2.weight.detach().float().to(dev), g.fc_t2.bias.detach().float().to(dev)))
        part = torch.stack([torch.nn.functional.linear(emb, blk.fc_t.weight.detach().float().to(dev),
                                                       blk.fc_t.bias.detach().float().to(dev)) for blk in blocks], 1)  # (T, L, 256)
        c2 = (b_res * math.sqrt(0.5)).unsqueeze(0).repeat(T, 1, 1)
        c2[:, :-1, :] += part[:, 1:, :]

        w0, b0 = self.init_conv[0].conv.folded()
        wf, bf = self.final_conv[0].conv.folded()
        out = getattr(self.final_conv, "2").conv
        if self.precision == "tf32":
            op = lambda w: round_to_tf32(w.float())
        else:
            op = lambda w: w.to(torch.bfloat16).contiguous()
        packed = {
            "w1": op(w1), "b1": b1.contiguous(),
            "w2": op(w2), "c2": c2.contiguous(),
            "part0": part[:, 0, :].contiguous(),
            "w0": w0.to(dev)[:, 0, 0].contiguous(), "b0": b0.to(dev).contiguous(),
            "ws": op(ws), "bs": bs.contiguous(),
            "wf": op(wf.to(dev)[:, :, 0]), "bf": bf.to(dev).contiguous(),
            "wo": out.weight.detach().float().to(dev)[0, :, 0].contiguous(),
            "bo": float(out.bias.detach().float()[0]),
        }
        return packed

    # ------------------------------------------------------------------------------------ forward --
    def forward(self, input_data):
        """WaveNet.py:164-172: ``(audio (B,1,L), diffusion_steps (B,1))`` -> eps (B,1,L).  Every call site of
        the reference passes one step for the whole batch (``t * ones((B,1))``, diffwave_ddpm.py:157,169,177);
        that is what the kernels implement, so mixed steps are rejected."""
        audio, diffusion_steps = input_data
        if torch.is_tensor(diffusion_steps):
            flat = diffusion_steps.reshape(-1)
            t = int(flat[0].item())
            if flat.numel() > 1 and not bool((flat == flat[0]).all()):
                raise NotImplementedError("per-sample diffusion steps are not used on the purification path")
        else:
            t = int(diffusion_steps)
        return self.engine().eps(audio, t)


class Engine:
    """Owns the C handle, the packed weights and the workspace for one device."""

    def __init__(self, model):
        lib = _lib.load()
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.AudioPureError(
                "audiopure_b200 runs on a B200 (sm_100a) only: move the model to a CUDA device. "
                "There is no CPU fallback.")
        self.device = dev
        self.lib = lib
        self.packed = model.pack_weights(dev)
        dc = model.diffusion_config
        self.hp = calc_diffusion_hyperparams(**dc)
        betas, _, acp = sde_tables(dc["T"], dc["beta_0"] * dc["T"], dc["beta_T"] * dc["T"])
        self._tables = [t.contiguous().float() for t in (self.hp["Alpha"], self.hp["Alpha_bar"], self.hp["Sigma"], betas, acp)]
        cfg = _lib.ApConfig()
        cfg.num_res_layers = model.num_res_layers
        cfg.dilation_cycle = model.dilation_cycle
        cfg.T = dc["T"]
        cfg.max_chunk = int(model.max_chunk)
        cfg.flags = _lib.AP_FLAG_TF32 if model.precision == "tf32" else 0
        for name, t in zip(("alpha", "alpha_bar", "sigma", "sde_beta", "sde_alphas_cumprod"), self._tables):
            setattr(cfg, name, ctypes.cast(t.data_ptr(), _lib.c_float_p))
        w = _lib.ApWeights()
        for k, v in self.packed.items():
            setattr(w, k, v if k == "bo" else v.data_ptr())
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.ap_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(self.handle)))
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ap_destroy(self.handle)
        except Exception:
            pass

    def precision(self):
        """-> ("bf16"|"tf32", round_bias): what the handle runs and what ap_create's tf32 probe found."""
        tf32, bias = ctypes.c_int(0), ctypes.c_uint32(0)
        _lib.check(self.lib.ap_precision(self.handle, ctypes.byref(tf32), ctypes.byref(bias)))
        return ("tf32" if tf32.value else "bf16"), bias.value

    def profile(self, enable):
        """Bracket every kernel launch with CUDA events (bench.py's roofline measurement)."""
        _lib.check(self.lib.ap_profile_enable(self.handle, 1 if enable else 0))

    def profile_read(self):
        """-> {"layer"|"tail"|"prologue": (device ms, launches)} accumulated since the last read."""
        ms = (ctypes.c_double * 3)(0.0, 0.0, 0.0)
        cnt = (ctypes.c_int64 * 3)(0, 0, 0)
        _lib.check(self.lib.ap_profile_read(self.handle, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(("layer", "tail", "prologue"))}

    def workspace(self, B, L):
        need = self.lib.ap_workspace_bytes(self.handle, B, L)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            assert self._ws.data_ptr() % 1024 == 0
        return self._ws

    def _prep(self, x):
        if not torch.is_tensor(x):
            x = torch.as_tensor(x)
        assert x.ndim == 3 and x.shape[1] == 1, "expected waveforms of shape (B, 1, L)"
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        return x, x.shape[0], x.shape[2]

    def _z(self, z, shape):
        if z is None:
            return None
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        assert tuple(z.shape) == tuple(shape), (tuple(z.shape), tuple(shape))
        return z

    def eps(self, x, t):
        x, B, L = self._prep(x)
        out = torch.empty_like(x)
        ws = self.workspace(B, L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ap_eps(self.handle, x.data_ptr(), B, L, int(t), out.data_ptr(), ws.data_ptr(),
                                       ws.numel(), _lib.stream_ptr()))
        return out

    def step(self, x, t, ca, cb, cc, z=None, seed=0, stream_id=0, clip_offset=0):
        x, B, L = self._prep(x)
        z = self._z(z, x.shape)
        out = torch.empty_like(x)
        ws = self.workspace(B, L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ap_step(self.handle, x.data_ptr(), out.data_ptr(), B, L, int(t), ca, cb, cc,
                                        z.data_ptr() if z is not None else None, seed, stream_id, clip_offset,
                                        ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out

    def _purify(self, fn, x, t, z, n_noise, seed, clip_offset):
        x, B, L = self._prep(x)
        z = self._z(z, (n_noise, B, 1, L))
        out = torch.empty_like(x)
        ws = self.workspace(B, L)
        with torch.cuda.device(self.device):
            _lib.check(fn(self.handle, x.data_ptr(), out.data_ptr(), B, L, int(t),
                          z.data_ptr() if z is not None else None, seed, clip_offset, ws.data_ptr(), ws.numel(),
                          _lib.stream_ptr()))
        return out

    def ddpm_purify(self, x, t_star, z=None, seed=0, clip_offset=0):
        return self._purify(self.lib.ap_ddpm_purify, x, t_star, z, t_star, seed, clip_offset)

    def sde_purify(self, x, t, z=None, seed=0, clip_offset=0):
        return self._purify(self.lib.ap_sde_purify, x, t, z, t + 1, seed, clip_offset)

    def one_shot(self, x, reverse_timestep):
        x, B, L = self._prep(x)
        out = torch.empty_like(x)
        ws = self.workspace(B, L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ap_one_shot(self.handle, x.data_ptr(), out.data_ptr(), B, L, int(reverse_timestep),
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out
