"""GPU: CosyVoice3LM (llm.py:664-705) through the same C ABI as the CosyVoice2 LM: the "llm" stage
recognises the variant by its state_dict (no llm_embedding, 6761-way bias-free head) and pads the head by 3 impossible ids.
Golden ids from the reference CosyVoice3LM.inference (tests/golden/lm3_l2*.npz).

First GPU run: round-1 driver test pass (green); the first-run xfail markers were removed in round 2."""
import pytest
import torch

from gpu_util import ctx
from oracle import cases, lm
from test_lm_gpu import _decode

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,cool", [("lm3_l2", 0.8), ("lm3_l2_stop", 1.0)])
def test_cosyvoice3lm_decode_ids_match_reference_fp32(tag, cool, golden):
    from cosyvoice_b200 import cvk
    g = golden(tag)
    c = cvk.Context(0, "fp32", workspace_gb=2.0)          # own context: the shared one holds the CosyVoice2 LM
    c.load_state_dict("llm", lm.synth_state_dict3(2, cool=cool), [2])
    assert c.lm_vocab() == 6764
    text, ptext, ptok, U = cases.lm3_case()
    ids = _decode(c, [text], [ptext], [ptok], U[:, None, :])[0]
    assert ids == g["ids"].tolist()


def test_cosyvoice3lm_bistream_ids_match_reference_fp32(golden):
    """§8 a7 for CosyVoice3LM (llm.py:551-661 as inherited by :664-705): fill 6564 / eos 6562, prompt text split at <|endofprompt|>,
    sos / task_id rows of speech_embedding behind cvk_lm_feed's kind 2 - ids identical to the reference's inference_bistream."""
    from cosyvoice_b200.model3 import B200CosyVoice3Model
    g = golden("lm3_bistream_l2")
    chunks, ptext, ptok, U = cases.bistream3_case()
    m = B200CosyVoice3Model(precision="fp32", device=0, workspace_gb=2.0)
    m.ctx.load_state_dict("llm", lm.bistream_state_dict3(2), [2])            # LM stage only
    ids = list(m.lm_generate_bistream(iter(chunks), ptext, ptok, uniforms=U))
    assert ids == g["ids"].tolist()
