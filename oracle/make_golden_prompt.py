"""Golden vectors for the prompt-side feature extractors (run in the build container; writes tests/golden/prompt_feat.npz).

kaldi fbank: the REAL torchaudio.compliance.kaldi.fbank with the frontend's arguments (cosyvoice/cli/frontend.py:108-113).
whisper log-mel: openai-whisper is not installed offline; the golden is produced by transformers' WhisperFeatureExtractor
(feature_size 128) - an independent implementation of whisper's published log_mel_spectrogram - on the UNPADDED waveform (the
reference calls whisper.log_mel_spectrogram directly, without the 30 s padding of the HF pipeline), and the restatement in
oracle/prompt_feat.py must agree with it.  The Slaney filterbank of oracle/mel.py (librosa restatement, used by the 24 kHz mel
frontend as well) is compared with transformers.audio_utils.mel_filter_bank on the way.

    python -m oracle.make_golden_prompt
"""
import os

import numpy as np
import torch

from . import mel as omel, prompt_feat as opf


def main():
    import torchaudio.compliance.kaldi as kaldi
    from transformers.audio_utils import mel_filter_bank
    from transformers.models.whisper.feature_extraction_whisper import WhisperFeatureExtractor
    out = {}
    waves = opf.test_waves()
    fe = WhisperFeatureExtractor(feature_size=128)
    for k, (n_fft, sr, n_mels, fmax) in enumerate(((400, 16000, 128, 8000.0), (1920, 24000, 80, 8000.0), (1920, 24000, 80, 12000.0))):
        ours = omel.librosa_mel_filterbank(sr, n_fft, n_mels=n_mels, fmin=0.0, fmax=fmax)
        hf = mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=fmax, sampling_rate=sr,
                             norm="slaney", mel_scale="slaney").T
        d = np.abs(ours - hf).max()
        print(f"slaney filterbank sr={sr} n_fft={n_fft} n_mels={n_mels} fmax={fmax}: max |oracle - transformers| = {d:.3g} (max weight {hf.max():.3g})")
        assert d < 1e-6 * max(1.0, hf.max()) + 1e-7
    for i, w in enumerate(waves):
        ref_k = kaldi.fbank(w[None], num_mel_bins=80, dither=0, sample_frequency=16000)
        ref_k = ref_k - ref_k.mean(dim=0, keepdim=True)
        mine_k = opf.kaldi_fbank(w[None])
        dk = (ref_k - mine_k).abs().max().item()
        ref_w = torch.from_numpy(fe._np_extract_fbank_features(w.numpy()[None].astype(np.float32), "cpu")[0]).float()
        mine_w = opf.whisper_log_mel(w)
        dw = (ref_w - mine_w).abs().max().item()
        print(f"utterance {i} ({w.numel()} samples): kaldi fbank [{tuple(ref_k.shape)}] max |torchaudio - oracle| = {dk:.3g}; "
              f"whisper log-mel [{tuple(ref_w.shape)}] max |transformers - oracle| = {dw:.3g}")
        assert ref_k.shape == mine_k.shape and dk < 2e-3
        assert ref_w.shape == mine_w.shape and dw < 2e-4
        out[f"wave{i}"] = w.numpy()
        out[f"kaldi{i}"] = ref_k.numpy()
        out[f"whisper{i}"] = ref_w.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prompt_feat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
