"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/refimport.py) on the seeded cases of oracle/cases.py with the synthetic weights of oracle/weights.py, and
check the oracle restatement against it while doing so.  Run in the build container only:

    python -m oracle.make_golden            # writes tests/golden, prints max|oracle - reference| per case
"""
import os
import sys
import types

import numpy as np
import torch

from . import cases, flow, hift, lm, mel, refimport, sampling, weights

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"  wrote {name}.npz", {k: tuple(np.asarray(v).shape) for k, v in arrs.items()})


class patched_rng:
    """Feed the reference's global-RNG draws from explicit tensors."""

    def __init__(self, noise, rand_ini):
        self.noise, self.rand_ini = noise, rand_ini

    def __enter__(self):
        self.o_randn_like, self.o_rand = torch.randn_like, torch.rand
        noise, rand_ini, o_randn_like = self.noise, self.rand_ini, self.o_randn_like
        torch.randn_like = lambda x, *a, **k: noise.clone() if x.shape == noise.shape else o_randn_like(x)
        torch.rand = lambda *a, **k: rand_ini.clone()

    def __exit__(self, *a):
        torch.randn_like, torch.rand = self.o_randn_like, self.o_rand


class patched_multinomial:
    """Tensor.multinomial := inverse CDF on explicit uniforms (oracle/sampling.py draw_index)."""

    def __init__(self, get_u):
        self.get_u = get_u

    def __enter__(self):
        self.orig = torch.Tensor.multinomial
        get_u = self.get_u

        def fake(t, n, replacement=False, generator=None):
            return torch.tensor([sampling.draw_index(t.detach().numpy(), get_u())])
        torch.Tensor.multinomial = fake

    def __exit__(self, *a):
        torch.Tensor.multinomial = self.orig


@torch.inference_mode()
def gen_hift():
    print("hift")
    ref = refimport.build_hift()
    shapes = hift.param_shapes()
    rsd = ref.state_dict()
    assert list(rsd.keys()) == list(shapes.keys())
    assert all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
    sd = weights.synth_state_dict(shapes, 1986, hift.SYNTH_GAINS)
    ref.load_state_dict(sd, strict=True)
    melx, noise, rand_ini = cases.hift_case()
    with patched_rng(noise, rand_ini):
        rwav, rs = ref.inference(speech_feat=melx)
    rf0 = ref.f0_predictor(melx)
    wav, s = hift.inference(sd, melx, noise, rand_ini)
    print("  max|oracle-ref| wav %.3g source %.3g f0 %.3g" % ((wav - rwav).abs().max(), (s - rs).abs().max(),
                                                              (hift.f0_predict(sd, melx) - rf0).abs().max()))
    # decode with the reference's own source injected (vocoder body only)
    rdec = ref.decode(x=melx, s=rs)
    save("hift_b2_t24", wav=rwav.numpy(), source=rs.numpy(), f0=rf0.numpy(), decode=rdec.numpy())
    # streaming glue: cache_source overwrite (generator.py:566-567)
    cache = rs[:1, :, :960].clone() * 0.5
    with patched_rng(noise[:1], rand_ini[:1]):
        rwav2, rs2 = ref.inference(speech_feat=melx[:1], cache_source=cache)
    save("hift_cache_source", wav=rwav2.numpy(), source=rs2.numpy(), cache=cache.numpy())


@torch.inference_mode()
def gen_hift_causal():
    """CosyVoice3 vocoder: cosyvoice/hifigan/generator.py:572-726 (CausalHiFTGenerator), offline and one streaming call."""
    print("hift_causal")
    from . import hift_causal as hc
    ref = refimport.build_hift_causal()
    shapes = hc.param_shapes()
    rsd = ref.state_dict()
    assert list(rsd.keys()) == list(shapes.keys())
    assert all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
    sd = weights.synth_state_dict(shapes, 1986, hc.SYNTH_GAINS)
    ref.load_state_dict(sd, strict=True)
    mel, rand_ini, sine_noise = cases.hift_causal_case()
    # the reference's constructor-time random tensors (not in the state_dict) are replaced by the case's tensors
    ref.m_source.l_sin_gen.rand_ini = rand_ini.clone()
    ref.m_source.l_sin_gen.sine_waves = sine_noise.clone()
    out = {}
    for name, finalize in (("final", True), ("chunk", False)):
        rwav, rs = ref.inference(speech_feat=mel, finalize=finalize)
        ref.f0_predictor.to(torch.float64)
        rf0 = ref.f0_predictor(mel.to(torch.float64), finalize=finalize).float()
        wav, s = hc.inference(sd, mel, rand_ini, sine_noise, finalize)
        print("  %s: max|oracle-ref| wav %.3g source %.3g f0 %.3g  (wav %s, |wav| max %.3g)" % (
            name, (wav - rwav).abs().max(), (s - rs).abs().max(), (hc.f0_predict(sd, mel, finalize) - rf0).abs().max(),
            tuple(rwav.shape), rwav.abs().max()))
        out.update({f"wav_{name}": rwav.numpy(), f"source_{name}": rs.numpy(), f"f0_{name}": rf0.numpy()})
    save("hift_causal", **out)


@torch.inference_mode()
def gen_flow():
    print("flow")
    for tag, kw in (("small", dict(enc_blocks=2, enc_up_blocks=1, num_mid_blocks=2, n_blocks=2)),
                    ("full", dict(enc_blocks=6, enc_up_blocks=4, num_mid_blocks=12, n_blocks=4))):
        cfg = flow.FlowCfg(**kw)
        ref = refimport.build_flow(**kw)
        shapes = flow.param_shapes(cfg)
        rsd = ref.state_dict()
        assert set(rsd.keys()) == set(shapes.keys())
        assert all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
        sd = weights.synth_state_dict(shapes, 1986, flow.SYNTH_GAINS)
        ref.load_state_dict(sd, strict=True)
        token, ptok, pfeat, emb = cases.flow_case()
        N, P = token.shape[1], ptok.shape[1]
        outs = {}
        for name, streaming, finalize in (("offline", False, True), ("stream_final", True, True), ("stream_chunk", True, False)):
            if tag == "full" and name == "stream_final":
                continue
            r, _ = ref.inference(token=token, token_len=torch.tensor([N]), prompt_token=ptok,
                                 prompt_token_len=torch.tensor([P]), prompt_feat=pfeat,
                                 prompt_feat_len=torch.tensor([2 * P]), embedding=emb, streaming=streaming, finalize=finalize)
            o = flow.inference(sd, token, ptok, pfeat, emb, cfg, streaming=streaming, finalize=finalize)
            print(f"  {tag}/{name}: max|oracle-ref| {(r - o).abs().max():.3g}")
            outs["mel_" + name] = r.numpy()
        # estimator alone, export_onnx.py recipe
        x, mask, mu, t, spks, cond = cases.estimator_case()
        for streaming in (False, True):
            r = ref.decoder.estimator(x, mask, mu, t, spks, cond, streaming=streaming)
            o = flow.estimator(sd, x, mask, mu, t, spks, cond, cfg, streaming)
            print(f"  {tag}/estimator streaming={streaming}: max|oracle-ref| {(r - o).abs().max():.3g}")
            outs["est_stream" if streaming else "est_offline"] = r.numpy()
        # encoder alone
        xe = torch.nn.functional.embedding(torch.cat([ptok, token], 1).long(), sd["input_embedding.weight"])
        r, _ = ref.encoder(xe, torch.tensor([N + P]), streaming=False)
        o = flow.encoder(sd, xe, cfg, False)
        print(f"  {tag}/encoder: max|oracle-ref| {(r - o).abs().max():.3g}")
        outs["enc_offline"] = r.numpy()
        save("flow_" + tag, **outs)
    z = flow.cfm_noise(8)
    save("cfm_noise", first8=z.numpy(), ref_first4=np.array([-1.1258, -1.1524, -0.2506, -0.4339], dtype=np.float32))


@torch.inference_mode()
def gen_dit():
    """CosyVoice3 flow: cosyvoice/flow/DiT/dit.py:145-176 (estimator) and cosyvoice/flow/flow.py:369-414 (inference)."""
    print("dit (CosyVoice3 flow)")
    from . import dit
    for tag, depth in (("small", 2), ("full", 22)):
        ref = refimport.build_flow3(depth)
        shapes = dit.flow_param_shapes(depth)
        rsd = ref.state_dict()
        assert set(rsd.keys()) == set(shapes.keys()), set(rsd.keys()) ^ set(shapes.keys())
        assert all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
        sd = weights.synth_state_dict(shapes, 1986, dit.SYNTH_GAINS)
        ref.load_state_dict(sd, strict=True)
        token, ptok, pfeat, emb = cases.flow_case()
        N, P = token.shape[1], ptok.shape[1]
        outs = {}
        for name, streaming, finalize in (("offline", False, True), ("stream_final", True, True), ("stream_chunk", True, False)):
            if tag == "full" and name == "stream_final":
                continue
            r, _ = ref.inference(token=token, token_len=torch.tensor([N]), prompt_token=ptok,
                                 prompt_token_len=torch.tensor([P]), prompt_feat=pfeat,
                                 prompt_feat_len=torch.tensor([2 * P]), embedding=emb, streaming=streaming, finalize=finalize)
            o = dit.inference(sd, token, ptok, pfeat, emb, depth, streaming=streaming, finalize=finalize)
            print(f"  {tag}/{name}: max|oracle-ref| {(r - o).abs().max():.3g}  (|mel| max {r.abs().max():.3g})")
            outs["mel_" + name] = r.numpy()
        x, mask, mu, t, spks, cond = cases.estimator_case(T=130)     # 130 frames: three 50-frame chunks in streaming mode
        for streaming in (False, True):
            r = ref.decoder.estimator(x, mask, mu, t, spks, cond, streaming=streaming)
            o = dit.estimator(sd, x, mask, mu, t, spks, cond, depth, streaming)
            print(f"  {tag}/estimator streaming={streaming}: max|oracle-ref| {(r - o).abs().max():.3g}  (|out| max {r.abs().max():.3g})")
            outs["est_stream" if streaming else "est_offline"] = r.numpy()
        save("dit_" + tag, **outs)


@torch.inference_mode()
def gen_lm():
    print("lm")
    for tag, NL in (("l2", 2), ("l24", 24)):
        ref = refimport.build_llm(num_layers=NL)
        sd = lm.synth_state_dict(NL)
        ref.load_state_dict(sd, strict=True)
        text, ptext, ptok, U = cases.lm_case()
        st = {"i": 0, "c": 0}

        def get_u():
            u = float(U[st["i"], min(st["c"], 1)])
            st["c"] += 1
            return u
        ids = []
        with patched_multinomial(get_u):
            gen = ref.inference(text=text, text_len=torch.tensor([text.shape[1]], dtype=torch.int32), prompt_text=ptext,
                                prompt_text_len=torch.tensor([ptext.shape[1]], dtype=torch.int32), prompt_speech_token=ptok,
                                prompt_speech_token_len=torch.tensor([ptok.shape[1]], dtype=torch.int32),
                                embedding=torch.zeros(0, 192))
            for tok in gen:
                ids.append(tok)
                st["i"] += 1
                st["c"] = 0
        o, logps = lm.inference(sd, text, ptext, ptok, U, NL, return_logp=True)
        print(f"  {tag}: {len(ids)} ids, oracle == reference: {o == ids}")
        # teacher-forced hidden state + log-probs from the reference HF model
        lm_in = lm.build_lm_input(sd, text, ptext, ptok)
        full = torch.cat([lm_in, torch.nn.functional.embedding(torch.tensor([ids[:16]]), sd["speech_embedding.weight"])], 1)
        hid = ref.llm.model(inputs_embeds=full, output_hidden_states=True, return_dict=True).hidden_states[-1]
        rlogp = ref.llm_decoder(hid[0]).log_softmax(-1)
        ohid, _ = lm.qwen2_forward(sd, full, None, NL)
        print(f"  {tag}: teacher-forced max|oracle-ref| hidden {(hid - ohid).abs().max():.3g}")
        save("lm_" + tag, ids=np.array(ids, dtype=np.int32), hidden_last4=hid[0, -4:].numpy(),
             logp_rows=rlogp[-4:, ::41].numpy(), logp_step0_top=np.sort(logps[0].numpy())[-32:])


def gen_lm3():
    """CosyVoice3LM.inference (llm.py:664-705 + inherited :458-549)."""
    print("lm (CosyVoice3LM)")
    NL = 2
    ref = refimport.build_llm3(num_layers=NL)
    for tag, cool in (("lm3_l2", 0.8), ("lm3_l2_stop", 1.0)):
        sd = lm.synth_state_dict3(NL, cool=cool)
        rsd = ref.state_dict()
        assert set(rsd.keys()) == set(sd.keys()), set(rsd.keys()) ^ set(sd.keys())
        ref.load_state_dict(sd, strict=True)
        text, ptext, ptok, U = cases.lm3_case()
        st = {"i": 0, "c": 0}

        def get_u():
            u = float(U[st["i"], min(st["c"], 1)])
            st["c"] += 1
            return u
        ids = []
        with patched_multinomial(get_u):
            for tok in ref.inference(text=text, text_len=torch.tensor([text.shape[1]], dtype=torch.int32), prompt_text=ptext,
                                     prompt_text_len=torch.tensor([ptext.shape[1]], dtype=torch.int32), prompt_speech_token=ptok,
                                     prompt_speech_token_len=torch.tensor([ptok.shape[1]], dtype=torch.int32), embedding=torch.zeros(0, 192)):
                ids.append(int(tok))
                st["i"] += 1
                st["c"] = 0
        o = lm.inference3(sd, text, ptext, ptok, U, NL)
        print(f"  {tag}: {len(ids)} ids, oracle == reference: {o == ids}")
        save(tag, ids=np.array(ids, dtype=np.int32))


def gen_bistream():
    """cosyvoice/llm/llm.py:551-661 (Qwen2LM.inference_bistream) with the text arriving in chunks."""
    print("lm bistream")
    NL = 2
    ref = refimport.build_llm(num_layers=NL)
    sd = lm.bistream_state_dict(NL)
    ref.load_state_dict(sd, strict=True)
    chunks, ptext, ptok, U = cases.bistream_case()
    st = {"i": 0, "c": 0}

    def get_u():
        u = float(U[st["i"], min(st["c"], 1)])
        st["c"] += 1
        return u
    orig = ref.sampling_ids

    def sampling_ids(weighted_scores, decoded_tokens, sampling_, ignore_eos=True):
        st["i"], st["c"] = len(decoded_tokens), 0          # uniforms are indexed by the position in out_tokens
        return orig(weighted_scores, decoded_tokens, sampling_, ignore_eos)
    ref.sampling_ids = sampling_ids
    ids = []
    with patched_multinomial(get_u):
        for tok in ref.inference_bistream(text=iter(chunks), prompt_text=ptext, prompt_text_len=torch.tensor([ptext.shape[1]], dtype=torch.int32),
                                          prompt_speech_token=ptok, prompt_speech_token_len=torch.tensor([ptok.shape[1]], dtype=torch.int32),
                                          embedding=torch.zeros(0, 192)):
            ids.append(int(tok))
    ref.sampling_ids = orig
    o, trace = lm.inference_bistream(sd, chunks, ptext, ptok, U, NL, return_trace=True)
    print(f"  {len(ids)} ids yielded, {sum(1 for t in trace if t == lm.FILL_TOKEN)} fill tokens, oracle == reference: {o == ids}")
    save("lm_bistream_l2", ids=np.array(ids, dtype=np.int32), trace=np.array(trace, dtype=np.int32))


def gen_bistream3():
    """cosyvoice/llm/llm.py:551-661 as inherited by CosyVoice3LM (:664-705): sos / task_id rows of speech_embedding, fill 6564,
    eos 6562, prompt text split at <|endofprompt|>."""
    print("lm bistream (CosyVoice3LM)")
    NL = 2
    ref = refimport.build_llm3(num_layers=NL)
    sd = lm.bistream_state_dict3(NL)
    ref.load_state_dict(sd, strict=True)
    chunks, ptext, ptok, U = cases.bistream3_case()
    st = {"i": 0, "c": 0}

    def get_u():
        u = float(U[st["i"], min(st["c"], 1)])
        st["c"] += 1
        return u
    orig = ref.sampling_ids

    def sampling_ids(weighted_scores, decoded_tokens, sampling_, ignore_eos=True):
        st["i"], st["c"] = len(decoded_tokens), 0          # uniforms are indexed by the position in out_tokens
        return orig(weighted_scores, decoded_tokens, sampling_, ignore_eos)
    ref.sampling_ids = sampling_ids
    ids = []
    with patched_multinomial(get_u):
        for tok in ref.inference_bistream(text=iter(chunks), prompt_text=ptext, prompt_text_len=torch.tensor([ptext.shape[1]], dtype=torch.int32),
                                          prompt_speech_token=ptok, prompt_speech_token_len=torch.tensor([ptok.shape[1]], dtype=torch.int32),
                                          embedding=torch.zeros(0, 192)):
            ids.append(int(tok))
    ref.sampling_ids = orig
    o, trace = lm.inference_bistream(sd, chunks, ptext, ptok, U, NL, return_trace=True, variant="cv3")
    print(f"  {len(ids)} ids yielded, {sum(1 for t in trace if t == lm.FILL3)} fill tokens, last {trace[-1]}, oracle == reference: {o == ids}")
    save("lm3_bistream_l2", ids=np.array(ids, dtype=np.int32), trace=np.array(trace, dtype=np.int32))


def gen_sampling():
    print("sampling")
    refimport.install()
    from cosyvoice.utils.common import ras_sampling
    logp, hist, U, ignore = cases.sampling_case()
    out, oo = [], []
    for i in range(logp.shape[0]):
        st = {"c": 0}

        def get_u():
            u = float(U[i, min(st["c"], 1)])
            st["c"] += 1
            return u
        sc = logp[i].clone()
        if ignore[i]:
            sc[6561] = -float("inf")                   # llm.py:157-158
        with patched_multinomial(get_u):
            out.append(ras_sampling(sc, hist[i].tolist(), 25, top_p=0.8, top_k=25, win_size=10, tau_r=0.1))
        oo.append(sampling.ras_sample(logp[i].numpy(), hist[i].tolist(), float(U[i, 0]), float(U[i, 1]), bool(ignore[i])))
    print("  oracle == reference:", out == oo, " fallback exercised:", sum(1 for i in range(len(out)) if out[i] != 0))
    save("sampling", ids=np.array(out, dtype=np.int32))


@torch.inference_mode()
def gen_mel():
    print("mel")
    refimport.install()
    from matcha.utils.audio import mel_spectrogram as ref_mel
    y = cases.mel_case()
    r = ref_mel(y, 1920, 80, 24000, 480, 1920, 0, 8000, center=False)
    o = mel.mel_spectrogram(y)
    print(f"  max|oracle-ref| {(r - o).abs().max():.3g}")
    save("mel_b2", mel=r.numpy())
    # CosyVoice3 feat_extractor (cosyvoice3.yaml:140-147: fmax null) on a length that is NOT a multiple of the hop: the reflect
    # padding is about the true last sample and floor(N / 480) frames come out
    y3 = cases.mel_case(B=2, n=24137, seed=6)
    r3 = ref_mel(y3, 1920, 80, 24000, 480, 1920, 0, None, center=False)
    o3 = mel.mel_spectrogram(y3, fmax=None)
    print(f"  fmax=None, N=24137: {tuple(r3.shape)}, max|oracle-ref| {(r3 - o3).abs().max():.3g}")
    save("mel_b2_cv3", mel=r3.numpy())


def gen_masks():
    print("masks")
    refimport.install()
    from cosyvoice.utils.mask import subsequent_chunk_mask
    m = subsequent_chunk_mask(4, 2)
    assert m.int().tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [1, 1, 1, 1], [1, 1, 1, 1]]   # mask.py:148-152
    m = subsequent_chunk_mask(130, 50)
    assert torch.equal(m, flow.chunk_attention_mask(130, 50))
    save("chunk_mask_130_50", mask=m.numpy())


def stream_noise(k, n_samples):
    """noise of the k-th vocoder call of a request (shared with tests/test_model_gpu.py)"""
    g = torch.Generator().manual_seed(7000 + k)
    return torch.randn(n_samples, 9, generator=g)


def gen_stream():
    """cosyvoice/cli/model.py:328-394 (CosyVoice2Model.tts) itself, stream=False and stream=True, on small modules."""
    print("stream (reference CosyVoice2Model.tts)")
    refimport.install()
    from cosyvoice.cli.model import CosyVoice2Model
    NL, kw = 2, dict(enc_blocks=2, enc_up_blocks=1, num_mid_blocks=2, n_blocks=2)
    llm = refimport.build_llm(num_layers=NL)
    llm.load_state_dict(lm.synth_state_dict(NL), strict=True)
    fl = refimport.build_flow(**kw)
    fl.load_state_dict(weights.synth_state_dict(flow.param_shapes(flow.FlowCfg(**kw)), 1986, flow.SYNTH_GAINS), strict=True)
    hf = refimport.build_hift()
    hf.load_state_dict(weights.synth_state_dict(hift.param_shapes(), 1986, hift.SYNTH_GAINS), strict=True)
    text, ptext, ptok, U = cases.lm_case()
    _, _, pfeat, emb = cases.flow_case(P=9)
    pfeat = pfeat[:, :18]
    out = {}
    for mode, stream in (("offline", False), ("stream", True)):
        model = CosyVoice2Model(llm, fl, hf, fp16=False)
        st = {"i": 0, "c": 0, "k": 0}

        def get_u():
            u = float(U[st["i"], min(st["c"], 1)])
            st["c"] += 1
            return u
        orig_mn, orig_rl = torch.Tensor.multinomial, torch.randn_like

        def fake_mn(t, n, replacement=False, generator=None):
            r = torch.tensor([sampling.draw_index(t.detach().numpy(), get_u())])
            return r

        def fake_rl(x, *a, **k):
            if x.dim() == 3 and x.shape[2] == 9:
                z = stream_noise(st["k"], x.shape[1]).unsqueeze(0)
                st["k"] += 1
                return z
            return orig_rl(x, *a, **k)
        # the LM thread consumes one (u1,u2) pair per generated token: advance the step counter from the token list length
        orig_infer = llm.sampling_ids

        def sampling_ids(weighted_scores, decoded_tokens, sampling_, ignore_eos=True):
            st["i"], st["c"] = len(decoded_tokens), 0
            return orig_infer(weighted_scores, decoded_tokens, sampling_, ignore_eos)
        llm.sampling_ids = sampling_ids
        torch.Tensor.multinomial, torch.randn_like = fake_mn, fake_rl
        try:
            chunks = [o["tts_speech"] for o in model.tts(text=text, flow_embedding=emb, llm_embedding=emb, prompt_text=ptext,
                                                         llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok, prompt_speech_feat=pfeat,
                                                         stream=stream)]
        finally:
            torch.Tensor.multinomial, torch.randn_like = orig_mn, orig_rl
            llm.sampling_ids = orig_infer
        print(f"  {mode}: {len(chunks)} chunks, lengths {[c.shape[1] for c in chunks]}, hop_len after = {model.token_hop_len}")
        out[mode + "_lens"] = np.array([c.shape[1] for c in chunks], dtype=np.int64)
        out[mode + "_wav"] = torch.cat(chunks, 1).numpy()
    save("stream_tts", **out)


def gen_stream3():
    """cosyvoice/cli/model.py:397-450 (CosyVoice3Model: inherited tts + its own token2wav, which re-runs the causal vocoder over
    the whole mel so far and emits the new samples), stream=False and stream=True, on small CosyVoice3 modules."""
    print("stream3 (reference CosyVoice3Model.tts)")
    refimport.install()
    from cosyvoice.cli.model import CosyVoice3Model
    from . import dit, hift_causal as hc
    NL, depth = 2, 2
    llm = refimport.build_llm3(num_layers=NL)
    llm.load_state_dict(lm.synth_state_dict3(NL), strict=True)
    fl = refimport.build_flow3(depth)
    fl.load_state_dict(weights.synth_state_dict(dit.flow_param_shapes(depth), 1986, dit.SYNTH_GAINS), strict=True)
    hf = refimport.build_hift_causal()
    hf.load_state_dict(weights.synth_state_dict(hc.param_shapes(), 1986, hc.SYNTH_GAINS), strict=True)
    text, ptext, ptok, U = cases.lm3_case()
    _, _, pfeat, emb = cases.flow_case(P=9)
    pfeat = pfeat[:, :18]
    _, rand_ini, sine_noise = cases.hift_causal_case(T=400)
    hf.m_source.l_sin_gen.rand_ini = rand_ini.clone()
    hf.m_source.l_sin_gen.sine_waves = sine_noise.clone()
    out = {}
    for mode, stream in (("offline", False), ("stream", True)):
        model = CosyVoice3Model(llm, fl, hf, fp16=False)
        st = {"i": 0, "c": 0}

        def get_u():
            u = float(U[st["i"], min(st["c"], 1)])
            st["c"] += 1
            return u
        orig_mn = torch.Tensor.multinomial

        def fake_mn(t, n, replacement=False, generator=None):
            return torch.tensor([sampling.draw_index(t.detach().numpy(), get_u())])
        orig_infer = llm.sampling_ids

        def sampling_ids(weighted_scores, decoded_tokens, sampling_, ignore_eos=True):
            st["i"], st["c"] = len(decoded_tokens), 0
            return orig_infer(weighted_scores, decoded_tokens, sampling_, ignore_eos)
        llm.sampling_ids = sampling_ids
        torch.Tensor.multinomial = fake_mn
        try:
            chunks = [o["tts_speech"] for o in model.tts(text=text, flow_embedding=emb, llm_embedding=emb, prompt_text=ptext,
                                                         llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok, prompt_speech_feat=pfeat,
                                                         stream=stream)]
        finally:
            torch.Tensor.multinomial = orig_mn
            llm.sampling_ids = orig_infer
        print(f"  {mode}: {len(chunks)} chunks, lengths {[c.shape[1] for c in chunks]}, hop_len after = {model.token_hop_len}")
        out[mode + "_lens"] = np.array([c.shape[1] for c in chunks], dtype=np.int64)
        out[mode + "_wav"] = torch.cat(chunks, 1).numpy()
    # the oracle pipeline reproduces the offline run
    ids = lm.inference3(lm.synth_state_dict3(NL), text, ptext, ptok, U, NL)
    mel = dit.inference(weights.synth_state_dict(dit.flow_param_shapes(depth), 1986, dit.SYNTH_GAINS),
                        torch.tensor([ids], dtype=torch.int32), ptok, pfeat, emb, depth)
    wav, _ = hc.inference(weights.synth_state_dict(hc.param_shapes(), 1986, hc.SYNTH_GAINS), mel, rand_ini, sine_noise, True)
    print(f"  oracle pipeline vs reference offline: max|d| {np.abs(wav.numpy() - out['offline_wav']).max():.3g}")
    save("stream3_tts", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["hift", "hift_causal", "flow", "dit", "lm", "lm3", "bistream", "bistream3", "sampling", "mel", "masks", "stream", "stream3"]
    for w in which:
        globals()["gen_" + w]()
