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
