"""CPU: host-side control flow of the product with the device primitives replaced by oracle arithmetic.

The text-streaming LM (llm.py:551-661) is mostly control flow (5:15 interleaving, fill-token forcing, replay quirk); the
product implements it in Python over four C-ABI primitives.  Here those primitives are faked with the CPU oracle, so the
host logic of `B200CosyVoice2Model.lm_generate_bistream` is checked against the reference's golden ids without a GPU.
(The same method over the real library is checked by tests/test_lm_gpu.py::test_bistream_ids_match_reference_fp32.)"""
import threading

import numpy as np
import torch

from oracle import cases, lm, sampling


class FakeLmContext:
    """cvk_lm_begin / cvk_lm_feed / cvk_lm_next_logp / cvk_ras_sample semantics on oracle.lm.qwen2_forward"""

    def __init__(self, sd, num_layers):
        self.sd, self.nl = sd, num_layers
        self.lock = threading.Lock()
        self.past, self.hidden = None, None
        self.fed = []

    def lm_session(self, B, ctx_len):
        return object()

    def lm_begin(self, sess, B=1):
        self.past, self.hidden = None, None

    def lm_feed(self, sess, ids, kinds):
        emb = []
        for i, k in zip(ids, kinds):
            table = {0: "llm.model.model.embed_tokens.weight", 1: "speech_embedding.weight", 2: "llm_embedding.weight"}[k]
            emb.append(self.sd[table][i])
        self.fed.append(len(ids))
        y, self.past = lm.qwen2_forward(self.sd, torch.stack(emb)[None], self.past, self.nl)
        self.hidden = y[:, -1]

    def lm_next_logp(self, sess, B=1):
        return lm.logprobs(self.sd, self.hidden)

    def ras_sample(self, logp, history, hist_count, uniforms, ignore_eos):
        hist = history[0, :int(hist_count[0])].tolist()
        top = sampling.ras_sample(logp[0].numpy(), hist, float(uniforms[0, 0]), float(uniforms[0, 1]), ignore_eos=bool(ignore_eos[0]))
        return torch.tensor([top], dtype=torch.int32)


def _model_with(ctx):
    from cosyvoice_b200.model import B200CosyVoice2Model
    m = object.__new__(B200CosyVoice2Model)          # no device: only the attributes the host logic touches
    m.ctx, m.stream, m.device = ctx, None, torch.device("cpu")
    m._sessions, m.uniforms_override, m.generator = {}, None, None
    m.silent_tokens = []
    return m


def test_bistream_host_logic_matches_reference(golden):
    g = golden("lm_bistream_l2")
    chunks, ptext, ptok, U = cases.bistream_case()
    ctx = FakeLmContext(lm.bistream_state_dict(2), 2)
    m = _model_with(ctx)
    ids = list(m.lm_generate_bistream(iter(chunks), ptext, ptok, uniforms=U))
    assert ids == g["ids"].tolist()
    # first model call: sos + (5 text, 15 speech) + (5 text, the 7 remaining prompt speech tokens)
    assert ctx.fed[0] == 1 + (5 + 15) + (5 + 7)
    # three calls push 5 positions: the two 5-text refills after a fill token, and the final phase, which replays the last
    # input (1 stale token embedding, llm.py:634-637) + the 3 leftover text ids + task_id (llm.py:643)
    assert ctx.fed.count(5) == 3
    assert all(n in (1, 5, 33) for n in ctx.fed)


def test_llm_job_generator_branch_collects_tokens(golden):
    """cli/model.py:113-128: generator text -> bi-stream decode -> tokens appended to the session list, end flag set"""
    g = golden("lm_bistream_l2")
    chunks, ptext, ptok, U = cases.bistream_case()
    m = _model_with(FakeLmContext(lm.bistream_state_dict(2), 2))
    m.uniforms_override = U[:, None, :]
    m.tts_speech_token_dict, m.llm_end_dict = {"u": []}, {"u": False}
    m.llm_job(iter(chunks), ptext, ptok, torch.zeros(0, 192), "u")
    assert m.tts_speech_token_dict["u"] == g["ids"].tolist() and m.llm_end_dict["u"] is True
