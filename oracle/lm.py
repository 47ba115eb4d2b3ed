"""Oracle (test infrastructure): CosyVoice2 speech-token LM restated in plain torch fp32 on CPU.

Follows cosyvoice/llm/llm.py: Qwen2LM.inference :458-502 (prompt assembly, min/max length),
inference_wrapper :536-549 (decode loop: step -> llm_decoder -> log_softmax -> sampling_ids -> stop test ->
speech_embedding of the sampled id), Qwen2Encoder.forward_one_step :242-254.

The transformer arithmetic is third-party: transformers ``Qwen2ForCausalLM`` (pinned 4.51.3 in requirements.txt,
NOT vendored; 5.5.0 installed in the build container).  Restated from modeling_qwen2.py:
RMSNorm eps 1e-6, q/k/v bias, no o/mlp bias, GQA 14/2 heads x 64, half-split RoPE theta 1e6, SwiGLU 4864, final
RMSNorm.  Model shape from the public Qwen2.5-0.5B config (CosyVoice-BlankEN/config.json is not in the repo).
Pinned against the reference Qwen2LM running on the installed transformers by oracle/make_golden.py.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import sampling

D, N_HEAD, N_KV, DH, D_FF = 896, 14, 2, 64, 4864
VOCAB_TEXT, SPEECH_TOKENS = 151936, 6561
V_OUT = SPEECH_TOKENS + 3
ROPE_THETA, RMS_EPS = 1.0e6, 1e-6
STOP_IDS = (6561, 6562, 6563)


def param_shapes(num_layers=24, with_lm_head=True):
    s = OrderedDict()
    s["llm_embedding.weight"] = (2, D)
    s["llm.model.model.embed_tokens.weight"] = (VOCAB_TEXT, D)
    for i in range(num_layers):
        p = f"llm.model.model.layers.{i}"
        s[p + ".self_attn.q_proj.weight"] = (D, D)
        s[p + ".self_attn.q_proj.bias"] = (D,)
        s[p + ".self_attn.k_proj.weight"] = (N_KV * DH, D)
        s[p + ".self_attn.k_proj.bias"] = (N_KV * DH,)
        s[p + ".self_attn.v_proj.weight"] = (N_KV * DH, D)
        s[p + ".self_attn.v_proj.bias"] = (N_KV * DH,)
        s[p + ".self_attn.o_proj.weight"] = (D, D)
        s[p + ".mlp.gate_proj.weight"] = (D_FF, D)
        s[p + ".mlp.up_proj.weight"] = (D_FF, D)
        s[p + ".mlp.down_proj.weight"] = (D, D_FF)
        s[p + ".input_layernorm.weight"] = (D,)
        s[p + ".post_attention_layernorm.weight"] = (D,)
    s["llm.model.model.norm.weight"] = (D,)
    if with_lm_head:
        s["llm.model.lm_head.weight"] = (VOCAB_TEXT, D)      # tied alias of embed_tokens, unused at inference
    s["llm_decoder.weight"] = (V_OUT, D)
    s["llm_decoder.bias"] = (V_OUT,)
    s["speech_embedding.weight"] = (V_OUT, D)
    return s


# embeddings ~N(0, 0.5^2) give unit-ish residual stream; a hot head makes the token distribution peaky enough
# for the nucleus cut and the repetition fallback to be exercised.
SYNTH_GAINS = {"llm_decoder.weight": 4.0}


def synth_state_dict(num_layers=24, seed=1986):
    from .weights import synth_state_dict as _s
    shapes = param_shapes(num_layers, with_lm_head=False)
    sd = _s(shapes, seed, SYNTH_GAINS)
    sd["llm.model.lm_head.weight"] = sd["llm.model.model.embed_tokens.weight"]
    return sd


def rmsnorm(x, w):
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + RMS_EPS))


def rope(x, pos):
    """x [B,H,L,64], pos [L] -> half-split rotate (modeling_qwen2.py rotate_half)."""
    inv = 1.0 / (ROPE_THETA ** (torch.arange(0, DH, 2, dtype=torch.float32) / DH))
    fr = pos.float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    x1, x2 = x[..., : DH // 2], x[..., DH // 2:]
    return x * cos + torch.cat([-x2, x1], dim=-1) * sin


def qwen2_forward(sd, x, past=None, num_layers=24):
    """x [1,L,896] new embeddings; past = list of (k,v) [1,2,Lp,64].  Returns (final-normed hidden [1,L,896], past)."""
    B, L, _ = x.shape
    Lp = 0 if past is None else past[0][0].shape[2]
    pos = torch.arange(Lp, Lp + L)
    new_past = []
    h = x
    for i in range(num_layers):
        p = f"llm.model.model.layers.{i}"
        xn = rmsnorm(h, sd[p + ".input_layernorm.weight"])
        q = F.linear(xn, sd[p + ".self_attn.q_proj.weight"], sd[p + ".self_attn.q_proj.bias"]).view(B, L, N_HEAD, DH).transpose(1, 2)
        k = F.linear(xn, sd[p + ".self_attn.k_proj.weight"], sd[p + ".self_attn.k_proj.bias"]).view(B, L, N_KV, DH).transpose(1, 2)
        v = F.linear(xn, sd[p + ".self_attn.v_proj.weight"], sd[p + ".self_attn.v_proj.bias"]).view(B, L, N_KV, DH).transpose(1, 2)
        q, k = rope(q, pos), rope(k, pos)
        if past is not None:
            k = torch.cat([past[i][0], k], dim=2)
            v = torch.cat([past[i][1], v], dim=2)
        new_past.append((k, v))
        kk = k.repeat_interleave(N_HEAD // N_KV, dim=1)
        vv = v.repeat_interleave(N_HEAD // N_KV, dim=1)
        s = torch.matmul(q, kk.transpose(-2, -1)) / math.sqrt(DH)
        causal = torch.arange(Lp + L)[None, :] > (pos[:, None])
        s = s.masked_fill(causal[None, None], float("-inf"))
        o = torch.matmul(torch.softmax(s.float(), dim=-1), vv).transpose(1, 2).reshape(B, L, D)
        h = h + F.linear(o, sd[p + ".self_attn.o_proj.weight"])
        xn = rmsnorm(h, sd[p + ".post_attention_layernorm.weight"])
        g = F.linear(xn, sd[p + ".mlp.gate_proj.weight"])
        u = F.linear(xn, sd[p + ".mlp.up_proj.weight"])
        h = h + F.linear(F.silu(g) * u, sd[p + ".mlp.down_proj.weight"])
    return rmsnorm(h, sd["llm.model.model.norm.weight"]), new_past


def build_lm_input(sd, text, prompt_text, prompt_speech_token):
    """llm.py:474-494: [sos, embed(prompt_text ++ text), task_id, speech_embedding(prompt tokens)] -> [1,L0,896]."""
    t = torch.cat([prompt_text, text], dim=1).long()
    text_emb = F.embedding(t, sd["llm.model.model.embed_tokens.weight"])
    sos = sd["llm_embedding.weight"][0].reshape(1, 1, -1)
    task = sd["llm_embedding.weight"][1].reshape(1, 1, -1)
    sp = F.embedding(prompt_speech_token.long(), sd["speech_embedding.weight"]) if prompt_speech_token.shape[1] else torch.zeros(1, 0, D)
    return torch.cat([sos, text_emb, task, sp], dim=1)


def length_bounds(text_len, min_ratio=2.0, max_ratio=20.0):
    """llm.py:497-498 (text_len = tts text length only, after the prompt length is subtracted)."""
    return int(text_len * min_ratio), int(text_len * max_ratio)


def logprobs(sd, hidden_last):
    """llm.py:542."""
    return F.log_softmax(F.linear(hidden_last, sd["llm_decoder.weight"], sd["llm_decoder.bias"]), dim=-1)


def inference(sd, text, prompt_text, prompt_speech_token, uniforms, num_layers=24, min_ratio=2.0, max_ratio=20.0,
              return_logp=False):
    """llm.py:458-549.  ``uniforms`` [max_len, 2] float32 draws (u1 nucleus, u2 fallback) per step.
    Returns the list of generated ids (stop id excluded) and optionally the per-step log-prob vectors."""
    lm_in = build_lm_input(sd, text, prompt_text, prompt_speech_token)
    min_len, max_len = length_bounds(text.shape[1], min_ratio, max_ratio)
    out, past, logps = [], None, []
    for i in range(max_len):
        y, past = qwen2_forward(sd, lm_in, past, num_layers)
        logp = logprobs(sd, y[:, -1]).squeeze(0)
        if return_logp:
            logps.append(logp.clone())
        top = sampling.ras_sample(logp.numpy(), out, float(uniforms[i, 0]), float(uniforms[i, 1]), ignore_eos=i < min_len)
        if top in STOP_IDS:
            break
        out.append(top)
        lm_in = sd["speech_embedding.weight"][top].reshape(1, 1, -1)
    return (out, logps) if return_logp else out


FILL_TOKEN = 6563          # llm.py:277 speech_token_size + 2
MIX_RATIO = (5, 15)        # llm.py:267


def bistream_state_dict(num_layers):
    """Synthetic weights for the text-streaming case: the plain synthetic model practically never emits the stop id, and
    inference_bistream has no length cap (llm.py:642-661 decodes 'until met eos'), so the eos logit is raised until eos is a likely
    draw once it is allowed (final phase only - it is masked during the interleaved phase, llm.py:627)."""
    sd = {k: v.clone() for k, v in synth_state_dict(num_layers).items()}
    sd["llm_decoder.bias"][6561] += 8.0
    sd["llm_decoder.bias"][6562] -= 30.0       # the other ids >= 6561 raise ValueError in the reference (llm.py:635, 653)
    sd["llm_decoder.bias"][6563] -= 30.0       # natural fill tokens are possible in principle; this case only has forced ones
    return sd


def inference_bistream(sd, text_chunks, prompt_text, prompt_speech_token, uniforms, num_layers=24, return_trace=False, variant="qwen2"):
    """llm.py:551-661 (Qwen2LM.inference_bistream, inherited by CosyVoice3LM).  text_chunks: list of int tensors [1,k] (the text
    generator); uniforms [n,2], row len(out_tokens) is consumed by the draw that produces out_tokens[len(out_tokens)].  Returns
    the yielded ids (fill tokens are kept in the internal history only) and optionally the list of out_tokens including fill
    tokens.  variant "qwen2": sos / task_id = llm_embedding rows 0 / 1, fill 6563, eos 6561 (llm.py:275-277); variant "cv3"
    (CosyVoice3LM, llm.py:681-684, 562-564, 583-588): sos / task_id = speech_embedding rows 6561 / 6563, fill 6564, eos 6562, head
    without bias, and the prompt text up to and including <|endofprompt|> (151646) is fed before the 5:15 interleaving starts."""
    emb_t = lambda ids: F.embedding(ids.long(), sd["llm.model.model.embed_tokens.weight"])
    if variant == "cv3":
        sos = sd["speech_embedding.weight"][SOS3].reshape(1, 1, -1)
        task = sd["speech_embedding.weight"][TASK3].reshape(1, 1, -1)
        fill_token, eos_token = FILL3, EOS3
        head = lambda y: F.log_softmax(F.linear(y, sd["llm_decoder.weight"]), dim=-1)
    else:
        sos = sd["llm_embedding.weight"][0].reshape(1, 1, -1)
        task = sd["llm_embedding.weight"][1].reshape(1, 1, -1)
        fill_token, eos_token = FILL_TOKEN, 6561
        head = lambda y: logprobs(sd, y)
    sp = F.embedding(prompt_speech_token.long(), sd["speech_embedding.weight"]) if prompt_speech_token.shape[1] else torch.zeros(1, 0, D)
    lm_input = sos
    out_tokens, yielded, past = [], [], None
    if variant == "cv3":
        flat = prompt_text.flatten().tolist()
        assert 151646 in flat, "<|endofprompt|> not detected in CosyVoice3 prompt_text"        # llm.py:585
        eop = flat.index(151646)
        lm_input = torch.cat([lm_input, emb_t(prompt_text[:, :eop + 1])], dim=1)
        prompt_text = prompt_text[:, eop + 1:]
    text_cache = emb_t(prompt_text)
    P = prompt_speech_token.shape[1]
    next_fill_index = (int(P / MIX_RATIO[1]) + 1) * MIX_RATIO[1] - P

    def step(lm_input, past, ignore_eos):
        y, past = qwen2_forward(sd, lm_input, past, num_layers)
        logp = head(y[:, -1]).squeeze(0)
        return logp, past

    for this_text in text_chunks:
        text_cache = torch.cat([text_cache, emb_t(this_text)], dim=1)
        while sp.shape[1] != 0:
            if text_cache.shape[1] >= MIX_RATIO[0]:
                lm_input = torch.cat([lm_input, text_cache[:, :MIX_RATIO[0]], sp[:, :MIX_RATIO[1]]], dim=1)
                text_cache, sp = text_cache[:, MIX_RATIO[0]:], sp[:, MIX_RATIO[1]:]
            else:
                break
        if sp.shape[1] == 0:
            if (len(out_tokens) != 0 and out_tokens[-1] == fill_token) or (len(out_tokens) == 0 and lm_input.shape[1] == 1):
                if text_cache.shape[1] >= MIX_RATIO[0]:
                    lm_input_text = text_cache[:, :MIX_RATIO[0]]
                    if len(out_tokens) != 0 and out_tokens[-1] == fill_token:
                        lm_input = lm_input_text
                    else:
                        lm_input = torch.cat([lm_input, lm_input_text], dim=1)
                    text_cache = text_cache[:, MIX_RATIO[0]:]
                else:
                    continue
            while True:
                logp, past = step(lm_input, past, True)
                if next_fill_index != -1 and len(out_tokens) == next_fill_index:
                    top = fill_token
                    next_fill_index += MIX_RATIO[1] + 1
                else:
                    i = len(out_tokens)
                    top = sampling.ras_sample(logp.numpy(), out_tokens, float(uniforms[i, 0]), float(uniforms[i, 1]), ignore_eos=True)
                if top == fill_token:
                    next_fill_index = len(out_tokens) + MIX_RATIO[1] + 1
                out_tokens.append(top)
                if top >= 6561:
                    if top == fill_token:
                        break
                    raise ValueError(f"should not get token {top}")
                yielded.append(top)
                lm_input = sd["speech_embedding.weight"][top].reshape(1, 1, -1)
    lm_input = torch.cat([lm_input, text_cache, task], dim=1)
    while True:
        logp, past = step(lm_input, past, False)
        i = len(out_tokens)
        top = sampling.ras_sample(logp.numpy(), out_tokens, float(uniforms[i, 0]), float(uniforms[i, 1]), ignore_eos=False)
        out_tokens.append(top)
        if top >= 6561:
            if top == eos_token:
                break
            raise ValueError(f"should not get token {top}")
        yielded.append(top)
        lm_input = sd["speech_embedding.weight"][top].reshape(1, 1, -1)
    return (yielded, out_tokens) if return_trace else yielded


# ------------------------------------------------------------------------------------------------ CosyVoice3LM (llm.py:664-705)
V_OUT3 = SPEECH_TOKENS + 200
SOS3, EOS3, TASK3, FILL3 = 6561, 6562, 6563, 6564
STOP_IDS3 = tuple(range(SPEECH_TOKENS, SPEECH_TOKENS + 200))


def param_shapes3(num_layers=24):
    """CosyVoice3LM: no llm_embedding (sos / task_id are rows 6561 / 6563 of speech_embedding), head 896 -> 6761 WITHOUT bias."""
    s = param_shapes(num_layers, with_lm_head=False)
    for k in ("llm_embedding.weight", "llm_decoder.bias"):
        s.pop(k)
    s["llm_decoder.weight"] = (V_OUT3, D)
    s["speech_embedding.weight"] = (V_OUT3, D)
    return s


def synth_state_dict3(num_layers=24, seed=1986, cool=0.8):
    from .weights import synth_state_dict as _s
    sd = _s(param_shapes3(num_layers), seed, SYNTH_GAINS)
    sd["llm.model.lm_head.weight"] = sd["llm.model.model.embed_tokens.weight"]
    # 200 of the 6761 output ids stop the decode (llm.py:704) and only one of them is masked before min_len: with an untrained
    # head the loop would end after ~10 tokens.  Cooling the stop rows keeps the synthetic utterance long enough to exercise
    # the repetition window (cool = 0.8: the golden runs to max_len); cool = 1.0 gives the second golden, which ends on a
    # sampled stop id before min_len.
    sd["llm_decoder.weight"][SPEECH_TOKENS + 1:] *= cool
    return sd


def bistream_state_dict3(num_layers, boost=2.6):
    """Synthetic CosyVoice3LM weights for the text-streaming case.  inference_bistream raises on every id >= 6561 except the fill
    token (interleaved phase) / the eos token (final phase) and has no length cap, and CosyVoice3LM's head has no bias to steer
    with.  So a constant channel is planted: column 0 of every embedding row is a large constant (it survives the residual stream
    and the final RMSNorm as a reliably positive feature), and column 0 of the head is zero for the speech ids, strongly negative
    for the special ids and positive for eos (6562) - an emulated bias.  Nothing masks eos in the interleaved phase
    (sampling_ids masks index 6561 only, llm.py:156-157 - the reference has the same hazard with a real model), so the boost is
    set where eos sits at the edge of the top-25 nucleus (a few percent per draw) and cases.bistream3_case picks a uniform stream
    on which the first eos draw falls in the final phase."""
    sd = {k: v.clone() for k, v in synth_state_dict3(num_layers, cool=0.0).items()}
    sd["llm.model.model.embed_tokens.weight"][:, 0] = 6.0
    sd["speech_embedding.weight"][:, 0] = 6.0
    sd["llm.model.lm_head.weight"] = sd["llm.model.model.embed_tokens.weight"]
    w = sd["llm_decoder.weight"]
    w[:, 0] = 0.0
    w[SPEECH_TOKENS:, 0] = -4.0
    w[EOS3, 0] = boost
    return sd


def inference3(sd, text, prompt_text, prompt_speech_token, uniforms, num_layers=24, min_ratio=2.0, max_ratio=20.0):
    """Qwen2LM.inference (llm.py:458-549) as inherited by CosyVoice3LM: prompt [speech_embedding[6561], embed(prompt_text ++
    text), speech_embedding[6563], speech_embedding(prompt tokens)]; ``ignore_eos`` masks index speech_token_size = 6561 exactly
    as the shared sampling_ids does (llm.py:156-157; in CosyVoice3 that index is the sos id, the eos id 6562 stays drawable);
    any id in 6561..6760 stops the loop (llm.py:544)."""
    t = torch.cat([prompt_text, text], dim=1).long()
    assert 151646 in t, "<|endofprompt|> not detected in CosyVoice3 text or prompt_text"       # llm.py:478-479
    text_emb = F.embedding(t, sd["llm.model.model.embed_tokens.weight"])
    sos = sd["speech_embedding.weight"][SOS3].reshape(1, 1, -1)
    task = sd["speech_embedding.weight"][TASK3].reshape(1, 1, -1)
    sp = F.embedding(prompt_speech_token.long(), sd["speech_embedding.weight"]) if prompt_speech_token.shape[1] else torch.zeros(1, 0, D)
    lm_in = torch.cat([sos, text_emb, task, sp], dim=1)
    min_len, max_len = length_bounds(text.shape[1], min_ratio, max_ratio)
    out, past = [], None
    for i in range(max_len):
        y, past = qwen2_forward(sd, lm_in, past, num_layers)
        logp = F.log_softmax(F.linear(y[:, -1], sd["llm_decoder.weight"]), dim=-1).squeeze(0)
        top = sampling.ras_sample(logp.numpy(), out, float(uniforms[i, 0]), float(uniforms[i, 1]), ignore_eos=i < min_len)
        if top in STOP_IDS3:
            break
        out.append(top)
        lm_in = sd["speech_embedding.weight"][top].reshape(1, 1, -1)
    return out
