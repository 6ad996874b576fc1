"""Oracle (test infrastructure): the runtime log-mel front-end, restated.

The reference builds its front-end from third-party torchaudio
(``adaptive_attack_eval.py:83-85``, ``certified_robustness_eval.py:81-83``):

    MelSpectrogram(n_fft=2048, hop_length=512, n_mels=32, norm='slaney',
                   pad_mode='constant', mel_scale='slaney')   # sample_rate=16000
    AmplitudeToDB(stype='power')                              # top_db=None

torchaudio (pinned 0.11.0 in requirements.txt:14; 2.11.0 in this image) is not
under /root/reference, so its published algorithm is restated here from
``torchaudio/functional/functional.py`` (spectrogram: 54-145, amplitude_to_DB:
390-404, melscale_fbanks: 425-520) without calling torchaudio, and pinned
against torchaudio itself by the golden fixture ``mel_*.npz``.
"""

import math

import numpy as np
import torch

SAMPLE_RATE = 16000
N_FFT = 2048
HOP = 512
N_MELS = 32
F_MIN = 0.0
F_MAX = SAMPLE_RATE / 2.0
AMIN = 1e-10


def _hz_to_mel_slaney(freq):
    f_sp = 200.0 / 3
    mels = (freq - 0.0) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - 0.0) / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        mels = min_log_mel + math.log(freq / min_log_hz) / logstep
    return mels


def _mel_to_hz_slaney(mels):
    f_sp = 200.0 / 3
    freqs = 0.0 + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - 0.0) / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def melscale_fbanks(n_freqs=N_FFT // 2 + 1, f_min=F_MIN, f_max=F_MAX, n_mels=N_MELS, sample_rate=SAMPLE_RATE):
    """functional.py:425-520 with mel_scale='slaney', norm='slaney' -> (n_freqs, n_mels) fp32."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_slaney(f_min)
    m_max = _hz_to_mel_slaney(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def power_spectrogram(x):
    """functional.py:54-145 as configured: zero 'constant' centre padding of n_fft//2,
    periodic Hann window, one-sided rFFT, |.|^2.  x: (B,1,L) -> (B,1,1025,frames)."""
    B, C, L = x.shape
    window = torch.hann_window(N_FFT, periodic=True)
    flat = x.reshape(-1, L)
    spec = torch.stft(flat, n_fft=N_FFT, hop_length=HOP, win_length=N_FFT, window=window, center=True,
                      pad_mode="constant", normalized=False, onesided=True, return_complex=True)
    spec = spec.reshape(B, C, spec.shape[-2], spec.shape[-1])
    return spec.abs().pow(2.0)


def log_mel(x, fb=None):
    """MelSpectrogram + AmplitudeToDB('power', top_db=None): (B,1,16000) -> (B,1,32,32) dB."""
    if fb is None:
        fb = melscale_fbanks()
    spec = power_spectrogram(x)
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    db = 10.0 * torch.log10(torch.clamp(mel, min=AMIN))
    db = db - 10.0 * math.log10(max(AMIN, 1.0))
    return db


def log_mel_f64(x):
    """Independent float64 numpy restatement (explicit framing + np.fft.rfft) used to
    bound the fp32 rounding of either implementation in the tests."""
    xn = x.detach().cpu().numpy().astype(np.float64)
    B, C, L = xn.shape
    pad = N_FFT // 2
    xp = np.pad(xn, ((0, 0), (0, 0), (pad, pad)))
    n_frames = 1 + L // HOP
    n = np.arange(N_FFT)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)
    frames = np.stack([xp[..., i * HOP:i * HOP + N_FFT] for i in range(n_frames)], axis=-2) * window
    power = np.abs(np.fft.rfft(frames, axis=-1)) ** 2  # (B,C,frames,1025)
    fb = melscale_fbanks().double().numpy()
    mel = power @ fb  # (B,C,frames,32)
    db = 10.0 * np.log10(np.maximum(mel, AMIN))
    return np.swapaxes(db, -1, -2)
