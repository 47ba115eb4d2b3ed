"""Oracle (test infrastructure): CosyVoice2 flow stage (token -> mel) restated in plain torch fp32 on CPU.

Follows, for ONE utterance (the reference asserts batch 1, flow/flow.py:246):
  * CausalMaskedDiffWithXvec.inference            cosyvoice/flow/flow.py:235-281
  * UpsampleConformerEncoder.forward              cosyvoice/transformer/upsample_encoder.py:244-307
    (LinearNoSubsampling subsampling.py:92-113, EspnetRelPositionalEncoding embedding.py:224-302,
     PreLookaheadLayer upsample_encoder.py:82-103, Upsample1D :59-63,
     ConformerEncoderLayer encoder_layer.py:160-236, RelPositionMultiHeadedAttention attention.py:225-330,
     PositionwiseFeedForward positionwise_feed_forward.py:47-56)
  * CausalConditionalCFM.forward / solve_euler    cosyvoice/flow/flow_matching.py:203-227, 71-124
  * CausalConditionalDecoder.forward              cosyvoice/flow/decoder.py:405-494
    (CausalConv1d :36-62, CausalBlock1D :65-78, Matcha ResnetBlock1D decoder.py:46-61,
     SinusoidalPosEmb :14-29, TimestepEmbedding :73-117, BasicTransformerBlock transformer.py:243-316)
  * masks                                         cosyvoice/utils/mask.py:127-158,161-236; common.py:188-196
Third-party arithmetic not vendored in the reference: diffusers==0.29.0 ``Attention`` (q/k/v no bias, out bias,
scale 1/sqrt(64), additive mask) and ``GELU`` (Linear + exact erf GELU) - restated, parity unpinned for those.
Hyper-parameters: examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:38-87.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

D_ENC, H_ENC, FF_ENC = 512, 8, 2048
C_EST, H_EST, DH_EST = 256, 8, 64
N_MEL = 80
CFG_RATE = 0.7
STATIC_CHUNK_TOK = 25           # cosyvoice2.yaml:16 chunk_size (tokens); x2 after the up-sampler; 50 mel frames


class FlowCfg:
    def __init__(self, enc_blocks=6, enc_up_blocks=4, num_mid_blocks=12, n_blocks=4):
        self.enc_blocks, self.enc_up_blocks = enc_blocks, enc_up_blocks
        self.num_mid_blocks, self.n_blocks = num_mid_blocks, n_blocks


def _enc_layer_shapes(s, p):
    s[p + ".self_attn.pos_bias_u"] = (H_ENC, D_ENC // H_ENC)
    s[p + ".self_attn.pos_bias_v"] = (H_ENC, D_ENC // H_ENC)
    for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
        s[f"{p}.self_attn.{n}.weight"] = (D_ENC, D_ENC)
        s[f"{p}.self_attn.{n}.bias"] = (D_ENC,)
    s[p + ".self_attn.linear_pos.weight"] = (D_ENC, D_ENC)
    s[p + ".feed_forward.w_1.weight"] = (FF_ENC, D_ENC)
    s[p + ".feed_forward.w_1.bias"] = (FF_ENC,)
    s[p + ".feed_forward.w_2.weight"] = (D_ENC, FF_ENC)
    s[p + ".feed_forward.w_2.bias"] = (D_ENC,)
    for n in ("norm_ff", "norm_mha"):
        s[f"{p}.{n}.weight"] = (D_ENC,)
        s[f"{p}.{n}.bias"] = (D_ENC,)


def _resnet_shapes(s, p, cin):
    s[p + ".mlp.1.weight"] = (C_EST, 4 * C_EST)
    s[p + ".mlp.1.bias"] = (C_EST,)
    for b, ci in (("block1", cin), ("block2", C_EST)):
        s[f"{p}.{b}.block.0.weight"] = (C_EST, ci, 3)
        s[f"{p}.{b}.block.0.bias"] = (C_EST,)
        s[f"{p}.{b}.block.2.weight"] = (C_EST,)
        s[f"{p}.{b}.block.2.bias"] = (C_EST,)
    s[p + ".res_conv.weight"] = (C_EST, cin, 1)
    s[p + ".res_conv.bias"] = (C_EST,)


def _tb_shapes(s, p):
    inner = H_EST * DH_EST
    s[p + ".norm1.weight"] = (C_EST,)
    s[p + ".norm1.bias"] = (C_EST,)
    for n in ("to_q", "to_k", "to_v"):
        s[f"{p}.attn1.{n}.weight"] = (inner, C_EST)
    s[p + ".attn1.to_out.0.weight"] = (C_EST, inner)
    s[p + ".attn1.to_out.0.bias"] = (C_EST,)
    s[p + ".norm3.weight"] = (C_EST,)
    s[p + ".norm3.bias"] = (C_EST,)
    s[p + ".ff.net.0.proj.weight"] = (4 * C_EST, C_EST)
    s[p + ".ff.net.0.proj.bias"] = (4 * C_EST,)
    s[p + ".ff.net.2.weight"] = (C_EST, 4 * C_EST)
    s[p + ".ff.net.2.bias"] = (C_EST,)


def param_shapes(cfg=None):
    """state_dict keys/shapes of the reference CausalMaskedDiffWithXvec."""
    cfg = cfg or FlowCfg()
    s = OrderedDict()
    s["input_embedding.weight"] = (6561, D_ENC)
    s["spk_embed_affine_layer.weight"] = (N_MEL, 192)
    s["spk_embed_affine_layer.bias"] = (N_MEL,)
    for e in ("embed", "up_embed"):
        pass
    s["encoder.embed.out.0.weight"] = (D_ENC, D_ENC)
    s["encoder.embed.out.0.bias"] = (D_ENC,)
    s["encoder.embed.out.1.weight"] = (D_ENC,)
    s["encoder.embed.out.1.bias"] = (D_ENC,)
    s["encoder.after_norm.weight"] = (D_ENC,)
    s["encoder.after_norm.bias"] = (D_ENC,)
    s["encoder.pre_lookahead_layer.conv1.weight"] = (D_ENC, D_ENC, 4)
    s["encoder.pre_lookahead_layer.conv1.bias"] = (D_ENC,)
    s["encoder.pre_lookahead_layer.conv2.weight"] = (D_ENC, D_ENC, 3)
    s["encoder.pre_lookahead_layer.conv2.bias"] = (D_ENC,)
    for i in range(cfg.enc_blocks):
        _enc_layer_shapes(s, f"encoder.encoders.{i}")
    s["encoder.up_layer.conv.weight"] = (D_ENC, D_ENC, 5)
    s["encoder.up_layer.conv.bias"] = (D_ENC,)
    s["encoder.up_embed.out.0.weight"] = (D_ENC, D_ENC)
    s["encoder.up_embed.out.0.bias"] = (D_ENC,)
    s["encoder.up_embed.out.1.weight"] = (D_ENC,)
    s["encoder.up_embed.out.1.bias"] = (D_ENC,)
    for i in range(cfg.enc_up_blocks):
        _enc_layer_shapes(s, f"encoder.up_encoders.{i}")
    s["encoder_proj.weight"] = (N_MEL, D_ENC)
    s["encoder_proj.bias"] = (N_MEL,)
    e = "decoder.estimator"
    s[e + ".time_mlp.linear_1.weight"] = (4 * C_EST, 320)
    s[e + ".time_mlp.linear_1.bias"] = (4 * C_EST,)
    s[e + ".time_mlp.linear_2.weight"] = (4 * C_EST, 4 * C_EST)
    s[e + ".time_mlp.linear_2.bias"] = (4 * C_EST,)
    _resnet_shapes(s, e + ".down_blocks.0.0", 320)
    for j in range(cfg.n_blocks):
        _tb_shapes(s, f"{e}.down_blocks.0.1.{j}")
    s[e + ".down_blocks.0.2.weight"] = (C_EST, C_EST, 3)
    s[e + ".down_blocks.0.2.bias"] = (C_EST,)
    for i in range(cfg.num_mid_blocks):
        _resnet_shapes(s, f"{e}.mid_blocks.{i}.0", C_EST)
        for j in range(cfg.n_blocks):
            _tb_shapes(s, f"{e}.mid_blocks.{i}.1.{j}")
    _resnet_shapes(s, e + ".up_blocks.0.0", 2 * C_EST)
    for j in range(cfg.n_blocks):
        _tb_shapes(s, f"{e}.up_blocks.0.1.{j}")
    s[e + ".up_blocks.0.2.weight"] = (C_EST, C_EST, 3)
    s[e + ".up_blocks.0.2.bias"] = (C_EST,)
    s[e + ".final_block.block.0.weight"] = (C_EST, C_EST, 3)
    s[e + ".final_block.block.0.bias"] = (C_EST,)
    s[e + ".final_block.block.2.weight"] = (C_EST,)
    s[e + ".final_block.block.2.bias"] = (C_EST,)
    s[e + ".final_proj.weight"] = (N_MEL, C_EST, 1)
    s[e + ".final_proj.bias"] = (N_MEL,)
    return s


SYNTH_GAINS = {}


# ----------------------------------------------------------------------------------------------- masks
def chunk_attention_mask(T, chunk):
    """mask.py:127-158 ``subsequent_chunk_mask`` (num_left_chunks = -1): key j visible from query i iff
    j < (i // chunk + 1) * chunk.  chunk <= 0 -> full attention.  Returns bool [T,T]."""
    if chunk <= 0:
        return torch.ones(T, T, dtype=torch.bool)
    pos = torch.arange(T)
    return pos.unsqueeze(0) < ((pos // chunk + 1) * chunk).unsqueeze(1)


def mask_to_bias(mask):
    """common.py:188-196."""
    return (1.0 - mask.float()) * -1.0e10


# --------------------------------------------------------------------------------------------- encoder
def rel_pos_table(T):
    """embedding.py:224-302: ESPnet relative table for a length-T input: rows m=0..2T-2 hold pe(r), r=T-1-m,
    pe(r)[2i]=sin(r*w_i), pe(r)[2i+1]=cos(r*w_i)."""
    r = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, D_ENC, 2, dtype=torch.float32) * -(math.log(10000.0) / D_ENC))
    pe = torch.zeros(2 * T - 1, D_ENC)
    pe[:, 0::2] = torch.sin(r * div)
    pe[:, 1::2] = torch.cos(r * div)
    return pe.unsqueeze(0)


def _embed(sd, p, x):
    """subsampling.py:92-113 + embedding.py:270-272: Linear -> LayerNorm(1e-5) -> * sqrt(d)."""
    x = F.linear(x, sd[p + ".out.0.weight"], sd[p + ".out.0.bias"])
    x = F.layer_norm(x, (D_ENC,), sd[p + ".out.1.weight"], sd[p + ".out.1.bias"], 1e-5)
    return x * math.sqrt(D_ENC)


def _rel_shift(x):
    """attention.py:225-247."""
    b, h, t, n = x.shape
    zero_pad = torch.zeros(b, h, t, 1)
    xp = torch.cat([zero_pad, x], dim=-1).view(b, h, n + 1, t)
    return xp[:, :, 1:].view_as(x)[:, :, :, : n // 2 + 1]


def _enc_layer(sd, p, x, mask, pos_emb):
    """encoder_layer.py:160-236 (normalize_before, no macaron, no conv module) + attention.py:249-330."""
    B, T, _ = x.shape
    dk = D_ENC // H_ENC
    res = x
    xn = F.layer_norm(x, (D_ENC,), sd[p + ".norm_mha.weight"], sd[p + ".norm_mha.bias"], 1e-12)
    a = p + ".self_attn"
    q = F.linear(xn, sd[a + ".linear_q.weight"], sd[a + ".linear_q.bias"]).view(B, T, H_ENC, dk)
    k = F.linear(xn, sd[a + ".linear_k.weight"], sd[a + ".linear_k.bias"]).view(B, T, H_ENC, dk).transpose(1, 2)
    v = F.linear(xn, sd[a + ".linear_v.weight"], sd[a + ".linear_v.bias"]).view(B, T, H_ENC, dk).transpose(1, 2)
    pp = F.linear(pos_emb, sd[a + ".linear_pos.weight"]).view(1, -1, H_ENC, dk).transpose(1, 2)
    qu = (q + sd[a + ".pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[a + ".pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = _rel_shift(torch.matmul(qv, pp.transpose(-2, -1)))
    scores = (ac + bd) / math.sqrt(dk)
    m = (~mask).unsqueeze(1)                               # [B,1,T,T] True where masked
    scores = scores.masked_fill(m, -float("inf"))
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    o = torch.matmul(attn, v).transpose(1, 2).reshape(B, T, D_ENC)
    x = res + F.linear(o, sd[a + ".linear_out.weight"], sd[a + ".linear_out.bias"])
    res = x
    xn = F.layer_norm(x, (D_ENC,), sd[p + ".norm_ff.weight"], sd[p + ".norm_ff.bias"], 1e-12)
    h = F.silu(F.linear(xn, sd[p + ".feed_forward.w_1.weight"], sd[p + ".feed_forward.w_1.bias"]))
    return res + F.linear(h, sd[p + ".feed_forward.w_2.weight"], sd[p + ".feed_forward.w_2.bias"])


def encoder(sd, x, cfg=None, streaming=False, context=None):
    """upsample_encoder.py:244-307.  x [1,T,512] (token embeddings), optional look-ahead context [1,3,512]
    -> [1,2T,512]."""
    cfg = cfg or FlowCfg()
    p = "encoder"
    T = x.shape[1]
    h = _embed(sd, p + ".embed", x)
    pos = rel_pos_table(T)
    mask = chunk_attention_mask(T, STATIC_CHUNK_TOK if streaming else 0).unsqueeze(0)
    # PreLookaheadLayer :82-103
    o = h.transpose(1, 2)
    if context is None or context.shape[1] == 0:
        o = F.pad(o, (0, 3))
    else:
        c = _embed(sd, p + ".embed", context)
        o = torch.cat([o, c.transpose(1, 2)], dim=2)
    o = F.leaky_relu(F.conv1d(o, sd[p + ".pre_lookahead_layer.conv1.weight"], sd[p + ".pre_lookahead_layer.conv1.bias"]))
    o = F.pad(o, (2, 0))
    o = F.conv1d(o, sd[p + ".pre_lookahead_layer.conv2.weight"], sd[p + ".pre_lookahead_layer.conv2.bias"])
    h = o.transpose(1, 2) + h
    for i in range(cfg.enc_blocks):
        h = _enc_layer(sd, f"{p}.encoders.{i}", h, mask, pos)
    # Upsample1D :59-63: nearest x2, left pad 4, conv k5
    o = h.transpose(1, 2).repeat_interleave(2, dim=2)
    o = F.pad(o, (4, 0))
    o = F.conv1d(o, sd[p + ".up_layer.conv.weight"], sd[p + ".up_layer.conv.bias"])
    h = o.transpose(1, 2)
    T2 = h.shape[1]
    h = _embed(sd, p + ".up_embed", h)
    pos = rel_pos_table(T2)
    mask = chunk_attention_mask(T2, 2 * STATIC_CHUNK_TOK if streaming else 0).unsqueeze(0)
    for i in range(cfg.enc_up_blocks):
        h = _enc_layer(sd, f"{p}.up_encoders.{i}", h, mask, pos)
    return F.layer_norm(h, (D_ENC,), sd[p + ".after_norm.weight"], sd[p + ".after_norm.bias"], 1e-5)


# ------------------------------------------------------------------------------------------- estimator
def _causal_block(sd, p, x, mask):
    """decoder.py:65-78: (x*mask) -> left-pad 2 -> conv k3 -> LN over channels -> Mish -> *mask.  x [B,C,T]."""
    h = F.conv1d(F.pad(x * mask, (2, 0)), sd[p + ".block.0.weight"], sd[p + ".block.0.bias"])
    h = F.layer_norm(h.transpose(1, 2), (C_EST,), sd[p + ".block.2.weight"], sd[p + ".block.2.bias"], 1e-5).transpose(1, 2)
    return F.mish(h) * mask


def _resnet(sd, p, x, mask, temb):
    """Matcha decoder.py:55-61 with CausalBlock1D."""
    h = _causal_block(sd, p + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"]).unsqueeze(-1)
    h = _causal_block(sd, p + ".block2", h, mask)
    return h + F.conv1d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])


def _tblock(sd, p, x, bias):
    """Matcha transformer.py:243-316 with diffusers Attention/GELU.  x [B,T,C], bias [B,T,T] additive."""
    B, T, _ = x.shape
    xn = F.layer_norm(x, (C_EST,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    q = F.linear(xn, sd[p + ".attn1.to_q.weight"]).view(B, T, H_EST, DH_EST).transpose(1, 2)
    k = F.linear(xn, sd[p + ".attn1.to_k.weight"]).view(B, T, H_EST, DH_EST).transpose(1, 2)
    v = F.linear(xn, sd[p + ".attn1.to_v.weight"]).view(B, T, H_EST, DH_EST).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) * (DH_EST ** -0.5) + bias.unsqueeze(1)
    o = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(B, T, H_EST * DH_EST)
    x = x + F.linear(o, sd[p + ".attn1.to_out.0.weight"], sd[p + ".attn1.to_out.0.bias"])
    xn = F.layer_norm(x, (C_EST,), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"], 1e-5)
    h = F.gelu(F.linear(xn, sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"]))
    return x + F.linear(h, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"])


def time_embedding(sd, t):
    """Matcha decoder.py:14-29 (dim 320, scale 1000) + TimestepEmbedding :103-117 (SiLU)."""
    e = "decoder.estimator"
    half = 160
    emb = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    emb = 1000 * t.unsqueeze(1) * emb.unsqueeze(0)
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = F.silu(F.linear(emb, sd[e + ".time_mlp.linear_1.weight"], sd[e + ".time_mlp.linear_1.bias"]))
    return F.linear(h, sd[e + ".time_mlp.linear_2.weight"], sd[e + ".time_mlp.linear_2.bias"])


def estimator(sd, x, mask, mu, t, spks, cond, cfg=None, streaming=False):
    """decoder.py:405-494.  x,mu,cond [B,80,T]; mask [B,1,T]; t [B]; spks [B,80] -> [B,80,T]."""
    cfg = cfg or FlowCfg()
    e = "decoder.estimator"
    B, _, T = x.shape
    temb = time_embedding(sd, t)
    h = torch.cat([x, mu, spks.unsqueeze(-1).expand(-1, -1, T), cond], dim=1)
    am = mask.bool().expand(B, T, T) if not streaming else (mask.bool() & chunk_attention_mask(T, 50).unsqueeze(0))
    bias = mask_to_bias(am)

    def stage(p, h):
        h = _resnet(sd, p + ".0", h, mask, temb)
        h = h.transpose(1, 2)
        for j in range(cfg.n_blocks):
            h = _tblock(sd, f"{p}.1.{j}", h, bias)
        return h.transpose(1, 2)

    h = stage(e + ".down_blocks.0", h)
    skip = h
    h = F.conv1d(F.pad(h * mask, (2, 0)), sd[e + ".down_blocks.0.2.weight"], sd[e + ".down_blocks.0.2.bias"])
    for i in range(cfg.num_mid_blocks):
        h = stage(f"{e}.mid_blocks.{i}", h)
    h = torch.cat([h, skip], dim=1)
    h = stage(e + ".up_blocks.0", h)
    h = F.conv1d(F.pad(h * mask, (2, 0)), sd[e + ".up_blocks.0.2.weight"], sd[e + ".up_blocks.0.2.bias"])
    h = _causal_block(sd, e + ".final_block", h, mask)
    out = F.conv1d(h * mask, sd[e + ".final_proj.weight"], sd[e + ".final_proj.bias"])
    return out * mask


def cfm_noise(T):
    """flow_matching.py:199-200: torch.manual_seed(0); randn(1,80,15000) - the fixed initial noise."""
    g = torch.Generator(device="cpu")
    g.manual_seed(0)
    return torch.randn([1, 80, 50 * 300], generator=g)[:, :, :T]


def cfm_solve(sd, mu, mask, spks, cond, n_timesteps=10, cfg=None, streaming=False, z=None, est=None):
    """flow_matching.py:203-227 + 71-124 (cosine schedule, Euler, classifier-free guidance 0.7).  ``est`` replaces the CosyVoice2
    U-Net estimator (the CosyVoice3 DiT, oracle/dit.py) with the same (x, mask, mu, t, spks, cond, streaming) signature."""
    T = mu.shape[2]
    x = cfm_noise(T) if z is None else z
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
    t, dt = t_span[0].unsqueeze(0), t_span[1] - t_span[0]
    zeros = torch.zeros_like(mu)
    for step in range(1, len(t_span)):
        x_in = torch.cat([x, x], 0)
        args = (x_in, torch.cat([mask, mask], 0), torch.cat([mu, zeros], 0), torch.cat([t, t], 0),
                torch.cat([spks, torch.zeros_like(spks)], 0), torch.cat([cond, zeros], 0))
        out = est(*args, streaming) if est is not None else estimator(sd, *args, cfg, streaming)
        d, dc = out[:1], out[1:]
        x = x + dt * ((1.0 + CFG_RATE) * d - CFG_RATE * dc)
        t = t + dt
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t
    return x.float()


def inference(sd, token, prompt_token, prompt_feat, embedding, cfg=None, n_timesteps=10, streaming=False,
              finalize=True, return_mu=False):
    """flow.py:235-281.  token [1,N] int, prompt_token [1,P] int, prompt_feat [1,Tp,80], embedding [1,192]
    -> mel [1,80,2(N+P)-Tp] (finalize=False: the last 3 tokens are look-ahead context only)."""
    cfg = cfg or FlowCfg()
    emb = F.linear(F.normalize(embedding, dim=1), sd["spk_embed_affine_layer.weight"], sd["spk_embed_affine_layer.bias"])
    tok = torch.cat([prompt_token, token], dim=1)
    x = F.embedding(torch.clamp(tok, min=0).long(), sd["input_embedding.weight"])
    if finalize:
        h = encoder(sd, x, cfg, streaming)
    else:
        h = encoder(sd, x[:, :-3], cfg, streaming, context=x[:, -3:])
    mel_len1 = prompt_feat.shape[1]
    mel_len2 = h.shape[1] - mel_len1
    mu = F.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"]).transpose(1, 2).contiguous()
    cond = torch.zeros(1, N_MEL, mel_len1 + mel_len2)
    cond[:, :, :mel_len1] = prompt_feat.transpose(1, 2)
    mask = torch.ones(1, 1, mel_len1 + mel_len2)
    if return_mu:
        return mu
    feat = cfm_solve(sd, mu, mask, emb, cond, n_timesteps, cfg, streaming)
    return feat[:, :, mel_len1:]
