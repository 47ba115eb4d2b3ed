"""GPU: B200CosyVoice3Model (cosyvoice/cli/model.py:397-450 + the inherited tts, :328-394) end to end through libcvk - the
CosyVoice3LM stage, the DiT flow (stage "flow3") and the causal vocoder (stage "hift3") - against the outputs of the reference's
own CosyVoice3Model.tts on the same synthetic modules (tests/golden/stream3_tts.npz, oracle/make_golden.py::gen_stream3), offline
and streaming: chunk schedule bit-exact, waveform to a measured bound.  BASELINE.json config #4's model family."""
import numpy as np
import pytest
import torch

from gpu_util import maxdiff
from oracle import cases, dit, hift_causal as hc, lm, weights

pytestmark = pytest.mark.gpu
_m = {}


def model():
    if "m" not in _m:
        from cosyvoice_b200.model3 import B200CosyVoice3Model
        m = B200CosyVoice3Model(precision="fp32", device=0, workspace_gb=4.0)
        _, rand_ini, sine_noise = cases.hift_causal_case(T=400)
        m.load_state_dicts(lm.synth_state_dict3(2), weights.synth_state_dict(dit.flow_param_shapes(2), 1986, dit.SYNTH_GAINS),
                           weights.synth_state_dict(hc.param_shapes(), 1986, hc.SYNTH_GAINS), rand_ini=rand_ini, sine_noise=sine_noise)
        _m["m"] = m
    return _m["m"]


@pytest.mark.parametrize("stream", [False, True])
def test_tts3_matches_reference_model(stream, golden):
    g = golden("stream3_tts")
    m = model()
    text, ptext, ptok, U = cases.lm3_case()
    _, _, pfeat, emb = cases.flow_case(P=9)
    m.uniforms_override = U[:, None, :]
    m.token_hop_len = 25
    try:
        chunks = [o["tts_speech"] for o in m.tts(text=text, flow_embedding=emb, llm_embedding=emb, prompt_text=ptext,
                                                 llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok,
                                                 prompt_speech_feat=pfeat[:, :18], stream=stream)]
    finally:
        m.uniforms_override = None
    key = "stream" if stream else "offline"
    assert [c.shape[1] for c in chunks] == g[key + "_lens"].tolist()          # chunk schedule / speech_offset bookkeeping: bit-exact
    wav = torch.cat(chunks, 1)
    ref = torch.from_numpy(g[key + "_wav"])
    d_head = maxdiff(wav[:, :24000], ref[:, :24000])
    rel = ((wav - ref).norm() / ref.norm()).item()
    print(f"[cv3 {key}] max|d| first second {d_head:.3g}, relative L2 over {wav.shape[1]} samples {rel:.3g}, max|d| {maxdiff(wav, ref):.3g}")
    assert d_head < 5e-3, d_head
    assert rel < 0.05, rel
    if stream:
        assert m.token_hop_len == 100


def test_tts3_with_text_generator_bistream(golden):
    """config #4's call pattern: text as a generator + stream=True on the CosyVoice3 stack - the LM thread runs the CosyVoice3LM
    text-streaming decode (fill 6564 / eos 6562 / <|endofprompt|> split; ids == the reference's inference_bistream golden) while
    the main thread streams DiT-flow + causal-vocoder chunks."""
    from cosyvoice_b200.model3 import B200CosyVoice3Model
    g = golden("lm3_bistream_l2")
    chunks, ptext, ptok, U = cases.bistream3_case()
    m = B200CosyVoice3Model(precision="fp32", device=0, workspace_gb=4.0)
    _, rand_ini, sine_noise = cases.hift_causal_case(T=400)
    m.load_state_dicts(lm.bistream_state_dict3(2), weights.synth_state_dict(dit.flow_param_shapes(2), 1986, dit.SYNTH_GAINS),
                       weights.synth_state_dict(hc.param_shapes(), 1986, hc.SYNTH_GAINS), rand_ini=rand_ini, sine_noise=sine_noise)
    _, _, pfeat, emb = cases.flow_case(P=9)
    m.uniforms_override = U[:, None, :]
    m.token_hop_len = 25
    m.silent_tokens = []          # the synthetic ids are uniform over the codebook: do not drop "silent" ones, count every id
    try:
        outs = [o["tts_speech"] for o in m.tts(text=iter(chunks), flow_embedding=emb, llm_embedding=emb, prompt_text=ptext,
                                               llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok[:, :9],
                                               prompt_speech_feat=pfeat[:, :18], stream=True)]
    finally:
        m.uniforms_override = None
    n_ids = len(g["ids"])
    assert sum(o.shape[1] for o in outs) == n_ids * 960
    assert len(outs) >= 2 and all(torch.isfinite(o).all() for o in outs)
