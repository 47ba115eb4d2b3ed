"""Synthetic CosyVoice2-0.5B weights and inputs for the benchmark / smoke paths (no checkpoints exist offline).

Random-init tensors of the reference architecture, keyed by the reference's state_dict names
(examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:23-111; shapes per SURVEY.md Appendix A), generated directly on
the GPU.  This module is product-side (bench.py / smoke use it); the oracle has its own CPU generator.
"""
import math
from collections import OrderedDict

import torch


# ---------------------------------------------------------------------------------------------- shapes
def llm_shapes(num_layers=24):
    s = OrderedDict()
    s["llm_embedding.weight"] = (2, 896)
    s["llm.model.model.embed_tokens.weight"] = (151936, 896)
    for i in range(num_layers):
        p = f"llm.model.model.layers.{i}"
        s[p + ".self_attn.q_proj.weight"] = (896, 896)
        s[p + ".self_attn.q_proj.bias"] = (896,)
        s[p + ".self_attn.k_proj.weight"] = (128, 896)
        s[p + ".self_attn.k_proj.bias"] = (128,)
        s[p + ".self_attn.v_proj.weight"] = (128, 896)
        s[p + ".self_attn.v_proj.bias"] = (128,)
        s[p + ".self_attn.o_proj.weight"] = (896, 896)
        s[p + ".mlp.gate_proj.weight"] = (4864, 896)
        s[p + ".mlp.up_proj.weight"] = (4864, 896)
        s[p + ".mlp.down_proj.weight"] = (896, 4864)
        s[p + ".input_layernorm.weight"] = (896,)
        s[p + ".post_attention_layernorm.weight"] = (896,)
    s["llm.model.model.norm.weight"] = (896,)
    s["llm_decoder.weight"] = (6564, 896)
    s["llm_decoder.bias"] = (6564,)
    s["speech_embedding.weight"] = (6564, 896)
    return s


def flow_shapes(enc_blocks=6, enc_up_blocks=4, num_mid_blocks=12, n_blocks=4):
    s = OrderedDict()
    s["input_embedding.weight"] = (6561, 512)
    s["spk_embed_affine_layer.weight"] = (80, 192)
    s["spk_embed_affine_layer.bias"] = (80,)

    def embed(p):
        s[p + ".out.0.weight"] = (512, 512)
        s[p + ".out.0.bias"] = (512,)
        s[p + ".out.1.weight"] = (512,)
        s[p + ".out.1.bias"] = (512,)

    def enc_layer(p):
        s[p + ".self_attn.pos_bias_u"] = (8, 64)
        s[p + ".self_attn.pos_bias_v"] = (8, 64)
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            s[f"{p}.self_attn.{n}.weight"] = (512, 512)
            s[f"{p}.self_attn.{n}.bias"] = (512,)
        s[p + ".self_attn.linear_pos.weight"] = (512, 512)
        s[p + ".feed_forward.w_1.weight"] = (2048, 512)
        s[p + ".feed_forward.w_1.bias"] = (2048,)
        s[p + ".feed_forward.w_2.weight"] = (512, 2048)
        s[p + ".feed_forward.w_2.bias"] = (512,)
        for n in ("norm_ff", "norm_mha"):
            s[f"{p}.{n}.weight"] = (512,)
            s[f"{p}.{n}.bias"] = (512,)

    embed("encoder.embed")
    s["encoder.after_norm.weight"] = (512,)
    s["encoder.after_norm.bias"] = (512,)
    s["encoder.pre_lookahead_layer.conv1.weight"] = (512, 512, 4)
    s["encoder.pre_lookahead_layer.conv1.bias"] = (512,)
    s["encoder.pre_lookahead_layer.conv2.weight"] = (512, 512, 3)
    s["encoder.pre_lookahead_layer.conv2.bias"] = (512,)
    for i in range(enc_blocks):
        enc_layer(f"encoder.encoders.{i}")
    s["encoder.up_layer.conv.weight"] = (512, 512, 5)
    s["encoder.up_layer.conv.bias"] = (512,)
    embed("encoder.up_embed")
    for i in range(enc_up_blocks):
        enc_layer(f"encoder.up_encoders.{i}")
    s["encoder_proj.weight"] = (80, 512)
    s["encoder_proj.bias"] = (80,)
    e = "decoder.estimator"
    s[e + ".time_mlp.linear_1.weight"] = (1024, 320)
    s[e + ".time_mlp.linear_1.bias"] = (1024,)
    s[e + ".time_mlp.linear_2.weight"] = (1024, 1024)
    s[e + ".time_mlp.linear_2.bias"] = (1024,)

    def resnet(p, cin):
        s[p + ".mlp.1.weight"] = (256, 1024)
        s[p + ".mlp.1.bias"] = (256,)
        for b, ci in (("block1", cin), ("block2", 256)):
            s[f"{p}.{b}.block.0.weight"] = (256, ci, 3)
            s[f"{p}.{b}.block.0.bias"] = (256,)
            s[f"{p}.{b}.block.2.weight"] = (256,)
            s[f"{p}.{b}.block.2.bias"] = (256,)
        s[p + ".res_conv.weight"] = (256, cin, 1)
        s[p + ".res_conv.bias"] = (256,)

    def tblock(p):
        s[p + ".norm1.weight"] = (256,)
        s[p + ".norm1.bias"] = (256,)
        for n in ("to_q", "to_k", "to_v"):
            s[f"{p}.attn1.{n}.weight"] = (512, 256)
        s[p + ".attn1.to_out.0.weight"] = (256, 512)
        s[p + ".attn1.to_out.0.bias"] = (256,)
        s[p + ".norm3.weight"] = (256,)
        s[p + ".norm3.bias"] = (256,)
        s[p + ".ff.net.0.proj.weight"] = (1024, 256)
        s[p + ".ff.net.0.proj.bias"] = (1024,)
        s[p + ".ff.net.2.weight"] = (256, 1024)
        s[p + ".ff.net.2.bias"] = (256,)

    def stage(p, cin):
        resnet(p + ".0", cin)
        for j in range(n_blocks):
            tblock(f"{p}.1.{j}")

    stage(e + ".down_blocks.0", 320)
    s[e + ".down_blocks.0.2.weight"] = (256, 256, 3)
    s[e + ".down_blocks.0.2.bias"] = (256,)
    for i in range(num_mid_blocks):
        stage(f"{e}.mid_blocks.{i}", 256)
    stage(e + ".up_blocks.0", 512)
    s[e + ".up_blocks.0.2.weight"] = (256, 256, 3)
    s[e + ".up_blocks.0.2.bias"] = (256,)
    s[e + ".final_block.block.0.weight"] = (256, 256, 3)
    s[e + ".final_block.block.0.bias"] = (256,)
    s[e + ".final_block.block.2.weight"] = (256,)
    s[e + ".final_block.block.2.bias"] = (256,)
    s[e + ".final_proj.weight"] = (80, 256, 1)
    s[e + ".final_proj.bias"] = (80,)
    return s


def hift_shapes():
    s = OrderedDict()

    def wn(prefix, wshape, transposed=False):
        s[prefix + ".bias"] = (wshape[1] if transposed else wshape[0],)
        s[prefix + ".parametrizations.weight.original0"] = (wshape[0], 1, 1)
        s[prefix + ".parametrizations.weight.original1"] = tuple(wshape)

    def resblock(prefix, ch, k):
        for i in range(3):
            wn(f"{prefix}.convs1.{i}", (ch, ch, k))
        for i in range(3):
            wn(f"{prefix}.convs2.{i}", (ch, ch, k))
        for i in range(3):
            s[f"{prefix}.activations1.{i}.alpha"] = (ch,)
        for i in range(3):
            s[f"{prefix}.activations2.{i}.alpha"] = (ch,)

    s["m_source.l_linear.weight"] = (1, 9)
    s["m_source.l_linear.bias"] = (1,)
    wn("conv_pre", (512, 80, 7))
    for i, k in enumerate((16, 11, 7)):
        wn(f"ups.{i}", (512 // 2 ** i, 512 // 2 ** (i + 1), k), transposed=True)
    for i, k in enumerate((30, 6, 1)):
        ch = 512 // 2 ** (i + 1)
        s[f"source_downs.{i}.weight"] = (ch, 18, k)
        s[f"source_downs.{i}.bias"] = (ch,)
    for i, k in enumerate((7, 7, 11)):
        resblock(f"source_resblocks.{i}", 512 // 2 ** (i + 1), k)
    for i in range(3):
        for j, k in enumerate((3, 7, 11)):
            resblock(f"resblocks.{i * 3 + j}", 512 // 2 ** (i + 1), k)
    wn("conv_post", (18, 64, 7))
    cin = 80
    for i in range(5):
        wn(f"f0_predictor.condnet.{2 * i}", (512, cin, 3))
        cin = 512
    s["f0_predictor.classifier.weight"] = (1, 512)
    s["f0_predictor.classifier.bias"] = (1,)
    return s


# ---------------------------------------------------------------------------------------------- CosyVoice3 shapes
def llm3_shapes(num_layers=24):
    """CosyVoice3LM (llm/llm.py:664-705): no llm_embedding, 6761-way head without bias, 6761-row speech embedding."""
    s = llm_shapes(num_layers)
    for k in ("llm_embedding.weight", "llm_decoder.bias"):
        s.pop(k)
    s["llm_decoder.weight"] = (6761, 896)
    s["speech_embedding.weight"] = (6761, 896)
    return s


def dit_flow_shapes(depth=22):
    """CausalMaskedDiffWithDiT (flow/flow.py:286-414, cosyvoice3.yaml: DiT dim 1024, depth 22, 16 heads x 64, ff_mult 2)."""
    s = OrderedDict()
    s["input_embedding.weight"] = (6561, 80)
    s["spk_embed_affine_layer.weight"] = (80, 192)
    s["spk_embed_affine_layer.bias"] = (80,)
    s["pre_lookahead_layer.conv1.weight"] = (1024, 80, 4)
    s["pre_lookahead_layer.conv1.bias"] = (1024,)
    s["pre_lookahead_layer.conv2.weight"] = (80, 1024, 3)
    s["pre_lookahead_layer.conv2.bias"] = (80,)
    p = "decoder.estimator."
    s[p + "time_embed.time_mlp.0.weight"] = (1024, 256)
    s[p + "time_embed.time_mlp.0.bias"] = (1024,)
    s[p + "time_embed.time_mlp.2.weight"] = (1024, 1024)
    s[p + "time_embed.time_mlp.2.bias"] = (1024,)
    s[p + "input_embed.proj.weight"] = (1024, 320)
    s[p + "input_embed.proj.bias"] = (1024,)
    for c in ("conv1", "conv2"):
        s[p + f"input_embed.conv_pos_embed.{c}.0.weight"] = (1024, 64, 31)
        s[p + f"input_embed.conv_pos_embed.{c}.0.bias"] = (1024,)
    for i in range(depth):
        b = p + f"transformer_blocks.{i}."
        s[b + "attn_norm.linear.weight"] = (6144, 1024)
        s[b + "attn_norm.linear.bias"] = (6144,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[b + f"attn.{n}.weight"] = (1024, 1024)
            s[b + f"attn.{n}.bias"] = (1024,)
        s[b + "ff.ff.0.0.weight"] = (2048, 1024)
        s[b + "ff.ff.0.0.bias"] = (2048,)
        s[b + "ff.ff.2.weight"] = (1024, 2048)
        s[b + "ff.ff.2.bias"] = (1024,)
    s[p + "norm_out.linear.weight"] = (2048, 1024)
    s[p + "norm_out.linear.bias"] = (2048,)
    s[p + "proj_out.weight"] = (80, 1024)
    s[p + "proj_out.bias"] = (80,)
    return s


def hift_causal_shapes():
    """CausalHiFTGenerator (hifigan/generator.py:572-726): conv_pre k5 (look-right 4), CausalConv1dUpsample = plain Conv1d weights,
    causal f0 predictor (first conv k4)."""
    s = hift_shapes()
    out = OrderedDict()
    for k, v in s.items():
        if k.startswith("conv_pre.parametrizations.weight.original1"):
            v = (512, 80, 5)
        if k.startswith("ups."):
            i = int(k.split(".")[1])
            cin, cout = 512 // 2 ** i, 512 // 2 ** (i + 1)
            kk = (16, 11, 7)[i]
            if k.endswith("bias"):
                v = (cout,)
            elif k.endswith("original0"):
                v = (cout, 1, 1)
            else:
                v = (cout, cin, kk)
        if k == "f0_predictor.condnet.0.parametrizations.weight.original1":
            v = (512, 80, 4)
        out[k] = v
    return out


DIT_GAINS = {"attn_norm.linear.weight": 0.1, "norm_out.linear.weight": 0.3, "conv_pos_embed": 0.5, "proj_out.weight": 0.012,
             "input_embed.proj.weight": 0.25}


def cosyvoice3_state_dicts(device, seed=1986, num_layers=24, depth=22):
    """(llm_sd, flow_sd, hift_sd) of Fun-CosyVoice3-0.5B shape (cosyvoice3.yaml) for bench config #4.  The head rows of the special
    ids are arranged like oracle.lm.bistream_state_dict3 (constant channel 0 + per-row weight = an emulated bias: CosyVoice3LM's head
    has none): no special id is ever drawn (the text-streaming decode raises on them, llm.py:635, 653); the benchmark ends the decode by a cap."""
    llm = random_state_dict(llm3_shapes(num_layers), device, seed, LLM_GAINS)
    llm["llm.model.model.embed_tokens.weight"][:, 0] = 6.0
    llm["speech_embedding.weight"][:, 0] = 6.0
    for i in range(num_layers):            # channel 0 of the residual stream stays the planted constant through every layer
        llm[f"llm.model.model.layers.{i}.self_attn.o_proj.weight"][0, :] = 0.0
        llm[f"llm.model.model.layers.{i}.mlp.down_proj.weight"][0, :] = 0.0
    w = llm["llm_decoder.weight"]
    w[:, 0] = 0.0
    w[6561:, 0] = -40.0                    # no special id is ever drawn; bench.py caps the decode (B200CosyVoice2Model.bistream_max_tokens)
    flow = random_state_dict(dit_flow_shapes(depth), device, seed + 1, DIT_GAINS)
    hift = random_state_dict(hift_causal_shapes(), device, seed + 2, HIFT_GAINS)
    return llm, flow, hift


def cv3_bistream_request(i, n_text=48, chunks=4):
    """config #4 request: the tts text arrives as a generator of `chunks` pieces (example.py:62-67 pattern), the prompt text
    contains <|endofprompt|> (151646, llm.py:585), 75 prompt speech tokens / 150 prompt mel frames."""
    u = z10_utterance(1000 + i, n_text)
    u["prompt_text"][0, 5] = 151646
    step = (n_text + chunks - 1) // chunks
    u["text_chunks"] = [u["text"][:, k:k + step] for k in range(0, n_text, step)]
    return u


# ---------------------------------------------------------------------------------------------- random init
def random_state_dict(shapes, device, seed, gains=None):
    """Scale rules keep activations O(1) through the depth of each stage (same rules as the test generator)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd = OrderedDict()
    for key, shape in shapes.items():
        gain = 1.0
        for pat, v in (gains or {}).items():
            if pat in key:
                gain *= v
        if key.endswith("parametrizations.weight.original0"):
            continue
        if key.endswith(".alpha"):
            t = 0.5 + torch.rand(shape, device=device, generator=g)
        elif "pos_bias_" in key:
            t = 0.1 * torch.randn(shape, device=device, generator=g)
        elif len(shape) == 1:
            t = 0.05 * gain * torch.randn(shape, device=device, generator=g) if key.endswith("bias") else \
                1.0 + 0.1 * torch.randn(shape, device=device, generator=g)
        elif "embed" in key and len(shape) == 2 and shape[0] > 512:
            t = 0.5 * gain * torch.randn(shape, device=device, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, device=device, generator=g) * (gain / math.sqrt(fan_in))
        sd[key] = t
    for key, shape in shapes.items():
        if key.endswith("parametrizations.weight.original0"):
            gain = 1.0
            for pat, v in (gains or {}).items():
                if pat in key:
                    gain *= v
            v = sd[key[:-1] + "1"]
            sd[key] = v.flatten(1).norm(dim=1).view(shape) * (1.0 + 0.1 * torch.randn(shape, device=device, generator=g)) * gain
    return OrderedDict((k, sd[k]) for k in shapes)


HIFT_GAINS = {"conv_post.parametrizations.weight.original0": 0.15, "f0_predictor.classifier.weight": 60.0,
              "f0_predictor.classifier.bias": 2000.0}
LLM_GAINS = {"llm_decoder.weight": 4.0}


def cosyvoice2_state_dicts(device, seed=1986, num_layers=24, flow_cfg=(6, 4, 12, 4), fixed_length=True):
    """(llm_sd, flow_sd, hift_sd) of CosyVoice2-0.5B shape.  fixed_length: bias the stop ids 6561..6563 to -1e4 so that
    a run with min_token_text_ratio == max_token_text_ratio produces exactly ratio * n_text tokens."""
    llm = random_state_dict(llm_shapes(num_layers), device, seed, LLM_GAINS)
    if fixed_length:
        llm["llm_decoder.bias"][6561:6564] = -1e4
    flow = random_state_dict(flow_shapes(*flow_cfg), device, seed + 1)
    hift = random_state_dict(hift_shapes(), device, seed + 2, HIFT_GAINS)
    return llm, flow, hift


def z10_utterance(i, n_text=50):
    """Canonical ~10 s zero-shot request: 12 prompt-text + n_text tts-text ids, 75 prompt speech tokens
    (3 s), prompt mel [1,150,80], speaker embedding [1,192]; utterance i uses seed 1986+i."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1986 + i)
    text = torch.randint(0, 151643, (1, n_text), dtype=torch.int32, generator=g)
    ptext = torch.randint(0, 151643, (1, 12), dtype=torch.int32, generator=g)
    ptok = torch.randint(0, 6561, (1, 75), dtype=torch.int32, generator=g)
    pfeat = torch.rand(1, 150, 80, generator=g) * 13.5 - 11.5
    emb = torch.randn(1, 192, generator=g)
    return dict(text=text, prompt_text=ptext, llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok,
                prompt_speech_feat=pfeat, llm_embedding=emb, flow_embedding=emb)


def batch32_zero_shot(batch=32, base=0, ragged=True):
    """BASELINE config #3: `batch` Z10 requests, tts text length ragged in {40..60} (=> 200..300 speech tokens at ratio 5)."""
    out = []
    for i in range(batch):
        n_text = 40 + ((base + i) * 7) % 21 if ragged else 50
        out.append(z10_utterance(base + i, n_text))
    return out
