"""CPU restatement of the CosyVoice3 flow stage: DiT estimator (cosyvoice/flow/DiT/dit.py:104-176, modules.py) and
CausalMaskedDiffWithDiT.inference (cosyvoice/flow/flow.py:369-414).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the
product never imports this file.

Pinned against the unmodified reference modules by oracle/make_golden.py (gen_dit) EXCEPT the rotary embedding: x_transformers
(requirements.txt:39, ==2.11.24) is not installed offline, so `rotary_freqs` / `apply_rotary` restate its published algorithm and
the reference is run with the same restatement stubbed in (oracle/refimport.py) - parity unpinned for that one piece.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import flow as _flow

DIM, HEADS, DH, FF, N_MEL = 1024, 16, 64, 2048, 80
CONV_K, CONV_GROUPS = 31, 16


def param_shapes(depth=22, prefix=""):
    """state_dict keys / shapes of the reference DiT (cosyvoice3.yaml: dim 1024, depth 22, heads 16 x 64, ff_mult 2)."""
    s = OrderedDict()
    p = prefix
    s[p + "time_embed.time_mlp.0.weight"] = (DIM, 256)
    s[p + "time_embed.time_mlp.0.bias"] = (DIM,)
    s[p + "time_embed.time_mlp.2.weight"] = (DIM, DIM)
    s[p + "time_embed.time_mlp.2.bias"] = (DIM,)
    s[p + "input_embed.proj.weight"] = (DIM, 4 * N_MEL)
    s[p + "input_embed.proj.bias"] = (DIM,)
    for c in ("conv1", "conv2"):
        s[p + f"input_embed.conv_pos_embed.{c}.0.weight"] = (DIM, DIM // CONV_GROUPS, CONV_K)
        s[p + f"input_embed.conv_pos_embed.{c}.0.bias"] = (DIM,)
    for i in range(depth):
        b = p + f"transformer_blocks.{i}."
        s[b + "attn_norm.linear.weight"] = (6 * DIM, DIM)
        s[b + "attn_norm.linear.bias"] = (6 * DIM,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[b + f"attn.{n}.weight"] = (DIM, DIM)
            s[b + f"attn.{n}.bias"] = (DIM,)
        s[b + "ff.ff.0.0.weight"] = (FF, DIM)
        s[b + "ff.ff.0.0.bias"] = (FF,)
        s[b + "ff.ff.2.weight"] = (DIM, FF)
        s[b + "ff.ff.2.bias"] = (DIM,)
    s[p + "norm_out.linear.weight"] = (2 * DIM, DIM)
    s[p + "norm_out.linear.bias"] = (2 * DIM,)
    s[p + "proj_out.weight"] = (N_MEL, DIM)
    s[p + "proj_out.bias"] = (N_MEL,)
    return s


def flow_param_shapes(depth=22):
    """CausalMaskedDiffWithDiT: token embedding 6561 x 80, speaker affine 192 -> 80, PreLookaheadLayer(80, 1024, 3), DiT."""
    s = OrderedDict()
    s["input_embedding.weight"] = (6561, N_MEL)
    s["spk_embed_affine_layer.weight"] = (N_MEL, 192)
    s["spk_embed_affine_layer.bias"] = (N_MEL,)
    s["pre_lookahead_layer.conv1.weight"] = (DIM, N_MEL, 4)
    s["pre_lookahead_layer.conv1.bias"] = (DIM,)
    s["pre_lookahead_layer.conv2.weight"] = (N_MEL, DIM, 3)
    s["pre_lookahead_layer.conv2.bias"] = (N_MEL,)
    s.update(param_shapes(depth, "decoder.estimator."))
    return s


# gains of the synthetic weights: AdaLN modulation / gates small enough that 22 blocks stay O(1), and a velocity field gentle
# enough (input / output projections) that ten Euler steps with CFG do not amplify fp32 summation-order noise (a random DiT
# at unit gain multiplies a 1e-5 difference by ~10 per step)
SYNTH_GAINS = {"attn_norm.linear.weight": 0.1, "norm_out.linear.weight": 0.3, "conv_pos_embed": 0.5, "proj_out.weight": 0.012,
               "input_embed.proj.weight": 0.25}


# ----------------------------------------------------------------------------------------------- rotary (x_transformers)
def rotary_freqs(T, dim=DH, base=10000.0):
    """RotaryEmbedding(dim_head).forward_from_seq_len(T): [1, T, dim], each frequency in two adjacent channels."""
    inv = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    f = torch.arange(T).float()[None, :, None] * inv[None, None, :]
    return torch.stack((f, f), dim=-1).flatten(-2)


def apply_rotary(t, freqs):
    """apply_rotary_pos_emb on the UN-SPLIT projection [B, T, heads*64]: only the first 64 channels (head 0) rotate
    (modules.py:368-373 passes the [b, n, inner_dim] tensor), pairs (2i, 2i+1) -> (x0 cos - x1 sin, x1 cos + x0 sin)."""
    rd = freqs.shape[-1]
    a, rest = t[..., :rd], t[..., rd:]
    a2 = a.reshape(*a.shape[:-1], -1, 2)
    rot = torch.stack((-a2[..., 1], a2[..., 0]), dim=-1).flatten(-2)
    return torch.cat((a * freqs.cos() + rot * freqs.sin(), rest), dim=-1)


# ----------------------------------------------------------------------------------------------- blocks
def time_embedding(sd, p, t):
    """modules.py:71-84, 606-616: sinus embedding (256, scale 1000, log(10000)/(half-1)) -> Linear -> SiLU -> Linear."""
    half = 128
    e = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    e = 1000.0 * t[:, None] * e[None]
    h = torch.cat((e.sin(), e.cos()), dim=-1)
    h = F.silu(F.linear(h, sd[p + "time_embed.time_mlp.0.weight"], sd[p + "time_embed.time_mlp.0.bias"]))
    return F.linear(h, sd[p + "time_embed.time_mlp.2.weight"], sd[p + "time_embed.time_mlp.2.bias"])


def conv_pos_embed(sd, p, x):
    """modules.py:115-145 CausalConvPositionEmbedding (mask None): two left-padded grouped Conv1d k31 g16 + Mish."""
    y = x.transpose(1, 2)
    for c in ("conv1", "conv2"):
        y = F.mish(F.conv1d(F.pad(y, (CONV_K - 1, 0)), sd[p + f"input_embed.conv_pos_embed.{c}.0.weight"],
                            sd[p + f"input_embed.conv_pos_embed.{c}.0.bias"], groups=CONV_GROUPS))
    return y.transpose(1, 2)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def dit_block(sd, b, x, t, mask, freqs):
    """modules.py:500-533 DiTBlock: AdaLN-zero modulation, attention (rotary on head 0 only, boolean mask), gated residuals,
    GELU(tanh) feed-forward."""
    B, T, _ = x.shape
    emb = F.linear(F.silu(t), sd[b + "attn_norm.linear.weight"], sd[b + "attn_norm.linear.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = torch.chunk(emb, 6, dim=1)
    h = _ln(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    q = apply_rotary(F.linear(h, sd[b + "attn.to_q.weight"], sd[b + "attn.to_q.bias"]), freqs)
    k = apply_rotary(F.linear(h, sd[b + "attn.to_k.weight"], sd[b + "attn.to_k.bias"]), freqs)
    v = F.linear(h, sd[b + "attn.to_v.weight"], sd[b + "attn.to_v.bias"])
    q, k, v = (z.view(B, T, HEADS, DH).transpose(1, 2) for z in (q, k, v))
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(DH)
    s = s.masked_fill(~mask, float("-inf"))
    o = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, T, HEADS * DH)
    o = F.linear(o, sd[b + "attn.to_out.0.weight"], sd[b + "attn.to_out.0.bias"])
    o = o.masked_fill(~mask[:, 0, -1].unsqueeze(-1), 0.0)          # modules.py:404-409 (4-D mask: last query row)
    x = x + gate_msa.unsqueeze(1) * o
    h = _ln(x) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
    h = F.gelu(F.linear(h, sd[b + "ff.ff.0.0.weight"], sd[b + "ff.ff.0.0.bias"]), approximate="tanh")
    return x + gate_mlp.unsqueeze(1) * F.linear(h, sd[b + "ff.ff.2.weight"], sd[b + "ff.ff.2.bias"])


def estimator(sd, x, mask, mu, t, spks, cond, depth=22, streaming=False, prefix="decoder.estimator.", chunk=50):
    """dit.py:145-176.  x, mu, cond [B,80,T]; mask [B,1,T]; t [B]; spks [B,80] -> [B,80,T]."""
    p = prefix
    B, _, T = x.shape
    te = time_embedding(sd, p, t)
    h = torch.cat([x.transpose(1, 2), cond.transpose(1, 2), mu.transpose(1, 2), spks[:, None, :].expand(B, T, spks.shape[1])], dim=-1)
    h = F.linear(h, sd[p + "input_embed.proj.weight"], sd[p + "input_embed.proj.bias"])
    h = conv_pos_embed(sd, p, h) + h
    freqs = rotary_freqs(T)
    pad = mask.bool()                                               # [B,1,T]
    if streaming:
        m = pad & _flow.chunk_attention_mask(T, chunk)[None]        # add_optional_chunk_mask(static chunk 50, all left chunks)
        m[m.sum(dim=-1) == 0] = True                                # mask.py:233-235
    else:
        m = pad.repeat(1, T, 1)
    m = m.unsqueeze(1)                                              # [B,1,T,T]
    for i in range(depth):
        h = dit_block(sd, p + f"transformer_blocks.{i}.", h, te, m, freqs)
    emb = F.linear(F.silu(te), sd[p + "norm_out.linear.weight"], sd[p + "norm_out.linear.bias"])
    scale, shift = torch.chunk(emb, 2, dim=1)
    h = _ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
    return F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"]).transpose(1, 2)


def pre_lookahead(sd, x, context=None, p="pre_lookahead_layer."):
    """upsample_encoder.py:66-103 with (in 80, channels 1024, look-ahead 3): conv k4 over [x | context or zeros] -> leaky_relu
    -> left-padded conv k3 -> + x."""
    y = x.transpose(1, 2)
    y = torch.cat([y, context.transpose(1, 2)], dim=2) if context is not None and context.shape[1] else F.pad(y, (0, 3))
    y = F.leaky_relu(F.conv1d(y, sd[p + "conv1.weight"], sd[p + "conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd[p + "conv2.weight"], sd[p + "conv2.bias"])
    return y.transpose(1, 2) + x


def inference(sd, token, prompt_token, prompt_feat, embedding, depth=22, n_timesteps=10, streaming=False, finalize=True,
              return_mu=False):
    """flow.py:369-414 CausalMaskedDiffWithDiT.inference.  token [1,N], prompt_token [1,P], prompt_feat [1,Tp,80],
    embedding [1,192] -> mel [1,80,2(N+P)-Tp] (finalize=False: the last 3 tokens are look-ahead context only)."""
    emb = F.linear(F.normalize(embedding, dim=1), sd["spk_embed_affine_layer.weight"], sd["spk_embed_affine_layer.bias"])
    tok = torch.cat([prompt_token, token], dim=1)
    x = F.embedding(torch.clamp(tok, min=0).long(), sd["input_embedding.weight"])
    h = pre_lookahead(sd, x) if finalize else pre_lookahead(sd, x[:, :-3], x[:, -3:])
    h = h.repeat_interleave(2, dim=1)
    mel_len1 = prompt_feat.shape[1]
    mel_len2 = h.shape[1] - mel_len1
    mu = h.transpose(1, 2).contiguous()
    if return_mu:
        return mu
    cond = torch.zeros(1, N_MEL, mel_len1 + mel_len2)
    cond[:, :, :mel_len1] = prompt_feat.transpose(1, 2)
    mask = torch.ones(1, 1, mel_len1 + mel_len2)
    est = lambda xx, mm, mmu, tt, ss, cc, st: estimator(sd, xx, mm, mmu, tt, ss, cc, depth, st)
    feat = _flow.cfm_solve(sd, mu, mask, emb, cond, n_timesteps, None, streaming, est=est)
    return feat[:, :, mel_len1:]
