"""Runtime log-mel front-end as ONE fused kernel: the ``transform`` slot of ``AcousticSystem``.

Stands in for the torchaudio composition every evaluation script of the reference builds
(``adaptive_attack_eval.py:83-85``, ``certified_robustness_eval.py:81-83``):

    Compose([MelSpectrogram(n_fft=2048, hop_length=512, n_mels=32, norm='slaney', pad_mode='constant',
                            mel_scale='slaney').cuda(), AmplitudeToDB(stype='power').cuda()])

``(B,1,L) fp32 -> (B,1,n_mels,1+L//512) fp32`` dB.  The slaney filterbank (host precompute, below)
follows torchaudio's ``melscale_fbanks``; it is stored sparsely (every mel filter is one run of
consecutive bins) for the kernel.
"""

import math

import numpy as np
import torch

from . import _lib

N_FFT = 2048
HOP = 512


def _hz_to_mel(freq):
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        return min_log_mel + math.log(freq / min_log_hz) / logstep
    return freq / f_sp


def _mel_to_hz(mels):
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    freqs = f_sp * mels
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def slaney_fbanks(n_mels=32, sample_rate=16000, f_min=0.0, f_max=None, n_freqs=N_FFT // 2 + 1):
    """(n_freqs, n_mels) fp32 triangular filters, mel_scale='slaney', norm='slaney'."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min), _hz_to_mel(f_max), n_mels + 2)
    f_pts = _mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


class _LogMelFn(torch.autograd.Function):
    """Forward: ap_logmel.  Backward: ap_logmel_backward (gradient w.r.t. the waveform only)."""

    @staticmethod
    def forward(ctx, x, module):
        ctx.module = module
        ctx.save_for_backward(x)
        return module._forward(x)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        return ctx.module._backward(x, grad_out), None


class LogMelSpectrogram(torch.nn.Module):
    """MelSpectrogram(n_fft=2048, hop=512, slaney/slaney, pad 'constant', power 2) + AmplitudeToDB('power')."""

    def __init__(self, n_mels=32, sample_rate=16000, f_min=0.0, f_max=None):
        super().__init__()
        self.n_mels = n_mels
        fb = slaney_fbanks(n_mels, sample_rate, f_min, f_max)
        self.register_buffer("fb", fb, persistent=False)
        starts, lens, offs, weights = [], [], [], []
        for m in range(n_mels):
            nz = torch.nonzero(fb[:, m]).flatten()
            if nz.numel() == 0:
                starts.append(0), lens.append(0), offs.append(len(weights))
                continue
            s, e = int(nz[0]), int(nz[-1]) + 1
            starts.append(s), lens.append(e - s), offs.append(len(weights))
            weights.extend(fb[s:e, m].tolist())
        k = np.arange(N_FFT // 2, dtype=np.float64)
        tw = np.stack([np.cos(2 * np.pi * k / N_FFT), -np.sin(2 * np.pi * k / N_FFT)], axis=1).astype(np.float32)
        self.register_buffer("twiddles", torch.from_numpy(tw), persistent=False)
        self.register_buffer("fb_start", torch.tensor(starts, dtype=torch.int32), persistent=False)
        self.register_buffer("fb_len", torch.tensor(lens, dtype=torch.int32), persistent=False)
        self.register_buffer("fb_off", torch.tensor(offs, dtype=torch.int32), persistent=False)
        self.register_buffer("fb_w", torch.tensor(weights, dtype=torch.float32), persistent=False)

    def _tables(self):
        return _lib.ApMelTables(self.twiddles.data_ptr(), self.fb_start.data_ptr(), self.fb_len.data_ptr(),
                                self.fb_off.data_ptr(), self.fb_w.data_ptr(), self.n_mels, int(self.fb_w.numel()))

    def forward(self, waveform):
        if waveform.device.type != "cuda":
            raise _lib.AudioPureError("LogMelSpectrogram runs on a CUDA device only (no CPU fallback)")
        if self.twiddles.device != waveform.device:
            self.to(waveform.device)
        assert waveform.ndim == 3 and waveform.shape[1] == 1, "expected (B, 1, L)"
        if torch.is_grad_enabled() and waveform.requires_grad:  # gradient-based attacks through AcousticSystem
            return _LogMelFn.apply(waveform, self)
        return self._forward(waveform)

    def _forward(self, waveform):
        lib = _lib.load()
        x = waveform.detach().to(torch.float32).contiguous()
        B, _, L = x.shape
        out = torch.empty(B, 1, self.n_mels, 1 + L // HOP, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.ap_logmel(x.data_ptr(), B, L, out.data_ptr(), self._tables(), _lib.stream_ptr()))
        return out

    def _backward(self, waveform, grad_out):
        lib = _lib.load()
        x = waveform.detach().to(torch.float32).contiguous()
        g = grad_out.to(torch.float32).contiguous()
        B, _, L = x.shape
        grad_x = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.ap_logmel_backward(x.data_ptr(), B, L, g.data_ptr(), grad_x.data_ptr(), self._tables(),
                                              _lib.stream_ptr()))
        return grad_x.to(waveform.dtype)
