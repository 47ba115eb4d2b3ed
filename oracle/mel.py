"""Oracle (test infrastructure): mel-spectrogram frontend restated on CPU.

Follows third_party/Matcha-TTS/matcha/utils/audio.py:45-82 (``mel_spectrogram``) with the
feat_extractor parameters of examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:150-158
(n_fft 1920, hop 480, win 1920, 80 mels, fmin 0, fmax 8000, center False).

The mel filterbank lives in librosa==0.10.2 (requirements.txt, NOT vendored and not installed
here): ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`` with its defaults htk=False
(Slaney scale) and norm='slaney'.  Restated below from the published algorithm and pinned
against the independent implementation in transformers.audio_utils.mel_filter_bank(norm="slaney",
mel_scale="slaney") by oracle/make_golden_prompt.py (max difference 2e-9 on the three filterbanks
in use: 24 kHz / 1920 / 80 with fmax 8000 and 12000, 16 kHz / 400 / 128); everything after the
filterbank is pinned against torch.stft through the reference function.
"""
import math

import numpy as np
import torch


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log_t = f >= min_log_hz
        mels = np.where(log_t, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    return np.where(log_t, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def librosa_mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney", dtype=np.float32):
    """librosa 0.10.2 ``filters.mel`` (Slaney scale, Slaney area normalisation)."""
    assert not htk and norm == "slaney"
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(dtype)


def mel_spectrogram(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920,
                    fmin=0, fmax=8000):
    """audio.py:45-82.  y [B, N] float32 -> [B, num_mels, N // hop]."""
    basis = torch.from_numpy(librosa_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)).float()
    window = torch.hann_window(win_size)
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    # explicit framed DFT (center=False): frames of win_size, hop hop_size
    n_frames = 1 + (y.shape[1] - n_fft) // hop_size
    idx = torch.arange(n_fft)[None, :] + hop_size * torch.arange(n_frames)[:, None]
    frames = y[:, idx] * window  # [B, F, n_fft]
    spec = torch.fft.rfft(frames.double(), n=n_fft, dim=-1)
    mag = torch.sqrt((spec.real ** 2 + spec.imag ** 2).float() + 1e-9).transpose(1, 2)  # [B, bins, F]
    mel = torch.matmul(basis, mag)
    return torch.log(torch.clamp(mel, min=1e-5))
