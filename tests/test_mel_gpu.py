"""GPU: mel-spectrogram frontend against the reference output (tests/golden/mel_b2.npz, matcha/utils/audio.py:45-82)."""
import numpy as np
import pytest
import torch

from gpu_util import ctx, maxdiff
from oracle import cases, mel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mel_golden(precision, golden):
    g = golden("mel_b2")
    y = cases.mel_case()
    c = ctx(precision)          # the frontend is fp32 in both modes
    out = c.mel_spectrogram(y.reshape(-1), [y.shape[1]] * 2).view(2, -1, 80).transpose(1, 2)
    d = maxdiff(out, torch.from_numpy(g["mel"]))
    assert d < 2e-3, d          # log-mel; fp32 DFT of 1920 points vs the reference's fp32 FFT


def test_mel_ragged():
    g = torch.Generator().manual_seed(3)
    ys = [torch.rand(1, n, generator=g) * 1.6 - 0.8 for n in (4800, 24000, 1920)]
    c = ctx("fp32")
    out = c.mel_spectrogram(torch.cat([y.reshape(-1) for y in ys]), [y.shape[1] for y in ys])
    o = 0
    for y in ys:
        ref = mel.mel_spectrogram(y)[0].t()
        assert maxdiff(out[o:o + ref.shape[0]], ref) < 2e-3
        o += ref.shape[0]


@pytest.mark.parametrize("precision", ["fp32"])
def test_mel_cv3_config_and_ragged_tail(precision, golden):
    """CosyVoice3 feat_extractor (fmax null = 12 kHz, cosyvoice3.yaml:140-147) on N = 24137 samples (N % 480 = 137): the reference
    reflect-pads about the TRUE last sample and emits floor(N / 480) frames (tests/golden/mel_b2_cv3.npz, made by the reference)."""
    g = golden("mel_b2_cv3")
    y = cases.mel_case(B=2, n=24137, seed=6)
    c = ctx(precision)
    out = c.mel_spectrogram(y.reshape(-1), [y.shape[1]] * 2, fmax=None).view(2, -1, 80).transpose(1, 2)
    assert out.shape == g["mel"].shape == (2, 80, 50)
    d = maxdiff(out, torch.from_numpy(g["mel"]))
    assert d < 2e-3, d
    from cosyvoice_b200.frontend import mel_spectrogram
    out2 = mel_spectrogram(y.cuda(), fmax=None, context=c)
    assert maxdiff(out2, torch.from_numpy(g["mel"])) < 2e-3
    with pytest.raises(Exception):
        c.mel_spectrogram(torch.zeros(700), [700])           # reflect padding by 720 needs more than 720 samples
