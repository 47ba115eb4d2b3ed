"""Oracle (test infrastructure): the two prompt-side feature extractors of cosyvoice/cli/frontend.py restated on the CPU.

* ``whisper_log_mel`` follows openai-whisper ``audio.py::log_mel_spectrogram`` as called at frontend.py:98
  (``whisper.log_mel_spectrogram(speech, n_mels=128)``): hann(400) periodic, torch.stft(n_fft 400, hop 160, center, reflect), power
  of all frames but the last, ``mel_filters(128)`` (= librosa.filters.mel(sr=16000, n_fft=400, n_mels=128), the package's
  assets/mel_filters.npz), log10(clamp 1e-10), floor at max - 8, (x + 4) / 4.  openai-whisper is a pip dependency of the reference
  (requirements.txt: openai-whisper==20231117) that is NOT installed here; the restatement is pinned against the independent
  implementation of the same published algorithm in ``transformers`` (WhisperFeatureExtractor, installed) by
  oracle/make_golden_prompt.py, which also pins the Slaney filterbank of oracle/mel.py against transformers' ``mel_filter_bank``.
* ``kaldi_fbank`` restates ``torchaudio.compliance.kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000)``
  (frontend.py:108-112; torchaudio IS installed: the goldens come from the real function) and the mean normalisation of :113.

Only tests/ and the golden generator import this module.
"""
import math

import numpy as np
import torch

from . import mel as omel


def whisper_log_mel(audio):
    """audio: float tensor [N] at 16 kHz -> [128, N // 160]"""
    window = torch.hann_window(400)
    stft = torch.stft(audio.float(), 400, 160, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    filters = torch.from_numpy(omel.librosa_mel_filterbank(16000, 400, n_mels=128, fmin=0.0, fmax=8000.0)).float()
    spec = filters @ mag
    log_spec = torch.clamp(spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def _kaldi_mel_banks(num_bins=80, n_fft=512, sr=16000.0, low=20.0, high=0.0):
    """torchaudio/compliance/kaldi.py::get_mel_banks without VTLN: [num_bins, n_fft/2] (+ a zero Nyquist column added by fbank)"""
    nbins = n_fft // 2
    nyq = 0.5 * sr
    if high <= 0.0:
        high += nyq
    width = sr / n_fft
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    lo, hi = mel(low), mel(high)
    delta = (hi - lo) / (num_bins + 1)
    b = np.arange(num_bins)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    m = mel(width * np.arange(nbins))[None, :]
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    return np.maximum(0.0, np.minimum(up, down))


def kaldi_fbank(wave, subtract_mean=True):
    """wave: float tensor [1, N] at 16 kHz -> [1 + (N - 400) // 160, 80] (float32 arithmetic like torchaudio)"""
    x = wave[0].float()
    n = x.numel()
    m = 1 + (n - 400) // 160
    frames = x.as_strided((m, 400), (160, 1)).clone()
    frames = frames - frames.mean(dim=1, keepdim=True)                       # remove_dc_offset
    prev = torch.cat([frames[:, :1], frames[:, :-1]], 1)                     # replicate-padded shift
    frames = frames - 0.97 * prev                                            # preemphasis_coefficient
    window = torch.hann_window(400, periodic=False).pow(0.85)                # povey
    frames = frames * window
    frames = torch.nn.functional.pad(frames, (0, 112))                       # round_to_power_of_two: 512
    power = torch.fft.rfft(frames).abs().pow(2.0)                            # [m, 257]
    banks = torch.from_numpy(_kaldi_mel_banks()).float()
    banks = torch.nn.functional.pad(banks, (0, 1))
    feat = torch.mm(power, banks.t())
    feat = torch.max(feat, torch.tensor(torch.finfo(torch.float).eps)).log()
    if subtract_mean:
        feat = feat - feat.mean(dim=0, keepdim=True)
    return feat


def test_waves():
    """three seeded 16 kHz utterances of different length: noise + a chirp, one with 0.2 s of digital silence in front"""
    g = torch.Generator().manual_seed(2024)
    out = []
    for n, silent in ((20800, 0), (7777, 0), (33000, 3200)):
        t = torch.arange(n) / 16000.0
        w = 0.3 * torch.sin(2 * math.pi * (200.0 + 1500.0 * t) * t) + 0.05 * torch.randn(n, generator=g)
        w[:silent] = 0.0
        out.append(w.clamp(-0.99, 0.99))
    return out
