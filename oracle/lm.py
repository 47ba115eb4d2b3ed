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
