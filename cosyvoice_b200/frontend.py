"""Drop-in for the reference's ``feat_extractor`` (matcha.utils.audio.mel_spectrogram bound in
examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:150-158, used by cosyvoice/cli/frontend.py:120-125)."""
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
