"""GPU: the prompt-side feature kernels (csrc/prompt_feat.cu) through the C ABI against tests/golden/prompt_feat.npz - kaldi fbank from
the real torchaudio.compliance.kaldi.fbank with the frontend's arguments (cosyvoice/cli/frontend.py:108-113), whisper log-mel from
transformers' WhisperFeatureExtractor on the unpadded waveform (frontend.py:98; openai-whisper itself is not installed offline, see
oracle/make_golden_prompt.py).  fp32 kernels: direct-DFT GEMM instead of an FFT, so the bound is round-off of a 400-term sum in
the log domain, not a model tolerance."""
import numpy as np
import pytest
import torch

from gpu_util import ctx, maxdiff

pytestmark = pytest.mark.gpu


def test_kaldi_fbank_and_whisper_log_mel_ragged_batch(golden):
    g = golden("prompt_feat")
    c = ctx("fp32")
    waves = [torch.from_numpy(g[f"wave{i}"]) for i in range(3)]
    lens = [w.numel() for w in waves]
    flat = torch.cat(waves)
    fb = c.kaldi_fbank(flat, lens, subtract_mean=True)
    lm = c.whisper_log_mel(flat, lens)
    o_k = o_w = 0
    for i, n in enumerate(lens):
        rk, rw = torch.from_numpy(g[f"kaldi{i}"]), torch.from_numpy(g[f"whisper{i}"]).t()
        mk, mw = rk.shape[0], rw.shape[0]
        assert mk == 1 + (n - 400) // 160 and mw == n // 160
        dk, dw = maxdiff(fb[o_k:o_k + mk], rk), maxdiff(lm[o_w:o_w + mw], rw)
        assert dk < 5e-3, (i, dk)          # log-energies spanning ~[-16, 6] minus their mean
        assert dw < 2e-3, (i, dw)          # (log10 + 4) / 4, values in ~[-1.5, 1.5]
        o_k, o_w = o_k + mk, o_w + mw
    assert o_k == fb.shape[0] and o_w == lm.shape[0]
    # one utterance alone == the same utterance inside the ragged batch (no cross-talk through the packed layout)
    alone = c.whisper_log_mel(waves[1], [lens[1]])
    assert torch.equal(alone, lm[lens[0] // 160:lens[0] // 160 + lens[1] // 160])
    # without the mean normalisation: torchaudio's raw fbank = golden + its frame mean is not stored, so check the relation instead
    raw = c.kaldi_fbank(waves[0], [lens[0]], subtract_mean=False)
    assert maxdiff(raw - raw.mean(dim=0, keepdim=True), fb[:raw.shape[0]]) < 1e-4


def test_frontend_wrappers_have_the_reference_layouts(golden):
    from cosyvoice_b200 import frontend
    g = golden("prompt_feat")
    w = torch.from_numpy(g["wave1"])
    m = frontend.log_mel_spectrogram(w[None].cuda())             # whisper: [1, 128, T] for a [1, N] input
    assert m.shape == (1, 128, w.numel() // 160)
    assert maxdiff(m[0], torch.from_numpy(g["whisper1"])) < 2e-3
    f = frontend.kaldi_fbank(w[None].cuda())
    assert f.shape == (1 + (w.numel() - 400) // 160, 80) and maxdiff(f, torch.from_numpy(g["kaldi1"])) < 5e-3
    with pytest.raises(ValueError):
        frontend.kaldi_fbank(w[None].cuda(), num_mel_bins=40)
    with pytest.raises(Exception):
        ctx("fp32").kaldi_fbank(w[:300], [300])                   # shorter than one 25 ms frame
