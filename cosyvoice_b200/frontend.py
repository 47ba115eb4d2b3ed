"""Drop-ins for the feature extractors of the reference frontend (cosyvoice/cli/frontend.py): ``mel_spectrogram`` for the
``feat_extractor`` (matcha.utils.audio.mel_spectrogram bound in examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:150-158, used at
frontend.py:120-125), ``log_mel_spectrogram`` for ``whisper.log_mel_spectrogram`` (frontend.py:98) and ``kaldi_fbank`` for
``torchaudio.compliance.kaldi.fbank`` + mean normalisation (frontend.py:108-113)."""
import torch

from . import cvk

_ctx = {}


def _context(device_index):
    if device_index not in _ctx:
        _ctx[device_index] = cvk.Context(device_index, "fp32", workspace_gb=1.0)
    return _ctx[device_index]


def mel_spectrogram(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920, fmin=0, fmax=8000, center=False,
                    context=None):
    """Same signature and result layout as matcha/utils/audio.py:45 ([B, num_mels, frames]).  libcvk implements the two
    feat_extractor configurations the reference ships - cosyvoice2.yaml:150-158 (fmax 8000) and cosyvoice3.yaml:140-147
    (fmax null = sr/2) - and raises for anything else (no silent fallback).  Any length N >= 721 is accepted: like the reference the
    signal is reflect-padded by 720 about its true first / last sample and floor(N / 480) frames are produced."""
    if (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, center) != (1920, 80, 24000, 480, 1920, 0, False) or fmax not in (8000, None, 12000):
        raise ValueError("libcvk implements the CosyVoice2 / CosyVoice3 feat_extractor configurations only (cosyvoice2.yaml:150-158, cosyvoice3.yaml:140-147)")
    if y.dim() != 2:
        raise ValueError("expected [B, N]")
    c = context or _context(y.device.index if y.is_cuda else 0)
    B, N = y.shape
    out = c.mel_spectrogram(y.reshape(-1), [N] * B, fmax=fmax)
    return out.view(B, N // hop_size, num_mels).transpose(1, 2).contiguous()


def log_mel_spectrogram(audio, n_mels=128, context=None):
    """whisper.log_mel_spectrogram(audio, n_mels=128) as the frontend calls it (frontend.py:98): audio [1, N] or [N] float at 16 kHz
    -> [1, 128, N // 160] (a 1-D input gives [128, T] like whisper).  Per-utterance dynamic-range floor, as for the reference's one
    utterance per call."""
    if n_mels != 128:
        raise ValueError("libcvk implements the 128-mel configuration the CosyVoice2/3 speech tokenizer uses")
    squeeze = audio.dim() == 1
    a = audio.reshape(1, -1) if squeeze else audio
    c = context or _context(a.device.index if a.is_cuda else 0)
    B, N = a.shape
    out = c.whisper_log_mel(a.reshape(-1), [N] * B).view(B, N // 160, 128).transpose(1, 2).contiguous()
    return out[0] if squeeze else out


def kaldi_fbank(waveform, num_mel_bins=80, dither=0, sample_frequency=16000, subtract_mean=True, context=None):
    """kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000) followed (subtract_mean=True) by the frontend's
    ``feat - feat.mean(dim=0, keepdim=True)`` (frontend.py:108-113): waveform [1, N] -> [1 + (N - 400) // 160, 80]."""
    if (num_mel_bins, dither, sample_frequency) != (80, 0, 16000) or waveform.dim() != 2 or waveform.shape[0] != 1:
        raise ValueError("libcvk implements the CAM++ front end of the reference only: [1, N] at 16 kHz, 80 bins, dither 0")
    c = context or _context(waveform.device.index if waveform.is_cuda else 0)
    return c.kaldi_fbank(waveform.reshape(-1), [waveform.shape[1]], subtract_mean=subtract_mean)
