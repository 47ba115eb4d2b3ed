"""Oracle (test infrastructure): CosyVoice3Model glue restated on the CPU oracles.

cosyvoice/cli/model.py:397-450 (CosyVoice3Model.__init__ / token2wav) + the inherited tts (:328-394): chunk schedule (hop 25
padded to the prompt length on the first chunk, doubled up to 100, 3 look-ahead tokens), the growing mel cache, the causal
vocoder re-run over the whole mel so far and the ``speech_offset`` bookkeeping.  The schedule does not depend on thread timing
(a chunk is cut only when enough tokens exist), so it is restated for an already finished LM.  Pinned against the reference's own
CosyVoice3Model.tts by oracle/make_golden.py::gen_stream3 (tests/golden/stream3_tts.npz).
"""
import math

import torch

from . import dit, hift_causal as hc

PRE_LOOKAHEAD = 3


def tts(ids, prompt_token, prompt_feat, embedding, flow_sd, hift_sd, depth, rand_ini, sine_noise, stream=False,
        hop=25, max_hop=100, scale=2):
    """Returns the list of waveform chunks [1,n_i] the reference yields."""
    tok_all = torch.tensor([ids], dtype=torch.int32)
    cache = {"mel": None, "speech_offset": 0}

    def token2wav(token, token_offset, streaming, finalize):
        mel = dit.inference(flow_sd, token, prompt_token, prompt_feat, embedding, depth, streaming=streaming, finalize=finalize)
        mel = mel[:, :, token_offset * 2:]
        if cache["mel"] is not None:
            mel = torch.cat([cache["mel"], mel], dim=2)
        cache["mel"] = mel
        wav, _ = hc.inference(hift_sd, mel, rand_ini, sine_noise, finalize)
        wav = wav[:, cache["speech_offset"]:]
        cache["speech_offset"] += wav.shape[1]
        return wav

    if not stream:
        return [token2wav(tok_all, 0, False, True)]
    out, offset = [], 0
    P = prompt_token.shape[1]
    pad = int(math.ceil(P / hop) * hop - P)
    while True:
        this_hop = hop + pad if offset == 0 else hop
        if len(ids) - offset >= this_hop + PRE_LOOKAHEAD:
            out.append(token2wav(tok_all[:, :offset + this_hop + PRE_LOOKAHEAD], offset, True, False))
            offset += this_hop
            hop = min(max_hop, hop * scale)
        else:
            break
    out.append(token2wav(tok_all, offset, False, True))       # the final call does not pass stream= (cli/model.py:366-373)
    return out
