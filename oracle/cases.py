"""Oracle (test infrastructure): the seeded synthetic inputs shared by the golden generator, the tests,
``smoke()`` and ``bench.py``.  Everything is derived from integer seeds so the GPU box can
rebuild inputs without the reference tree."""
import torch


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def hift_case(B=2, T=24, seed=0):
    g = _g(1000 + seed)
    mel = torch.randn(B, 80, T, generator=g) * 2 - 5
    noise = torch.randn(B, T * 480, 9, generator=g)
    rand_ini = torch.rand(B, 9, generator=g)
    rand_ini[:, 0] = 0
    return mel, noise, rand_ini


def hift_causal_case(T=24, seed=5):
    """CosyVoice3 vocoder case: mel [1,80,T] and the 'stored noise' of the causal source module (SineGen2.rand_ini [1,9] with
    column 0 zero, SineGen2.sine_waves [1,480T,9] uniform like the reference's torch.rand, generator.py:223-226)."""
    g = _g(1000 + seed)
    mel = torch.randn(1, 80, T, generator=g) * 2 - 5
    rand_ini = torch.rand(1, 9, generator=g)
    rand_ini[:, 0] = 0
    sine_noise = torch.rand(1, T * 480, 9, generator=g)
    return mel, rand_ini, sine_noise


def flow_case(N=31, P=20, seed=1):
    g = _g(2000 + seed)
    token = torch.randint(0, 6561, (1, N), dtype=torch.int32, generator=g)
    ptok = torch.randint(0, 6561, (1, P), dtype=torch.int32, generator=g)
    pfeat = torch.rand(1, 2 * P, 80, generator=g) * 13.5 - 11.5        # log-mel range, log(1e-5) = -11.5
    emb = torch.randn(1, 192, generator=g)
    return token, ptok, pfeat, emb


def estimator_case(T=64, seed=2):
    """export_onnx.py:34-41 recipe: x, mu, cond ~ U[0,1) [2,80,T], mask = 1, t ~ U[0,1) [2], spks [2,80]."""
    g = _g(3000 + seed)
    x = torch.rand(2, 80, T, generator=g)
    mask = torch.ones(2, 1, T)
    mu = torch.rand(2, 80, T, generator=g)
    t = torch.rand(2, generator=g)
    spks = torch.rand(2, 80, generator=g)
    cond = torch.rand(2, 80, T, generator=g)
    return x, mask, mu, t, spks, cond


def lm_case(n_text=7, n_prompt_text=4, n_prompt_speech=9, seed=3, n_uniform=400):
    g = _g(4000 + seed)
    text = torch.randint(0, 151643, (1, n_text), dtype=torch.int32, generator=g)
    ptext = torch.randint(0, 151643, (1, n_prompt_text), dtype=torch.int32, generator=g)
    ptok = torch.randint(0, 6561, (1, n_prompt_speech), dtype=torch.int32, generator=g)
    U = torch.rand(n_uniform, 2, generator=g)
    return text, ptext, ptok, U


def lm3_case():
    """CosyVoice3LM case: lm_case with the <|endofprompt|> id (151646, llm.py:478-479) closing the prompt text."""
    text, ptext, ptok, U = lm_case(seed=7)
    ptext = ptext.clone()
    ptext[0, -1] = 151646
    return text, ptext, ptok, U


def bistream_case(seed=11, n_uniform=400):
    """Text-streaming LM case: prompt text 6 ids, prompt speech 22 tokens (one full 5:15 block + 7 left over, first forced fill
    after 8 generated tokens), text arriving in chunks of 3 / 4 / 2 / 5 / 3 ids."""
    g = _g(4000 + seed)
    ptext = torch.randint(0, 151643, (1, 6), dtype=torch.int32, generator=g)
    ptok = torch.randint(0, 6561, (1, 22), dtype=torch.int32, generator=g)
    chunks = [torch.randint(0, 151643, (1, n), dtype=torch.int32, generator=g) for n in (3, 4, 2, 5, 3)]
    U = torch.rand(n_uniform, 2, generator=g)
    return chunks, ptext, ptok, U


def bistream3_case():
    """CosyVoice3LM text-streaming case (llm.py:583-588): the bistream case with <|endofprompt|> (151646) as the third prompt-text
    id - ids 0..2 are fed ahead of the 5:15 interleaving, ids 3..5 start the text cache.  Uniform stream chosen (seed 13) so that,
    with oracle.lm.bistream_state_dict3(2, boost=2.6), no special id is drawn in the interleaved phase and eos (6562) ends the
    final phase after 71 yielded ids and 3 forced fill tokens."""
    chunks, ptext, ptok, _ = bistream_case()
    _, _, _, U = bistream_case(seed=13)
    ptext = ptext.clone()
    ptext[0, 2] = 151646
    return chunks, ptext, ptok, U


def sampling_case(n=64, V=6564, seed=4):
    """Random log-prob vectors of varying peakiness + random decoded histories + uniforms."""
    g = _g(5000 + seed)
    temps = torch.rand(n, 1, generator=g) * 6 + 0.5
    logits = torch.randn(n, V, generator=g) * temps
    logp = torch.log_softmax(logits, dim=-1)
    hist = torch.randint(0, V, (n, 12), generator=g)
    # make repetitions likely: plant the arg-max in half of the histories
    am = logp.argmax(dim=-1)
    for i in range(0, n, 2):
        hist[i, -3] = am[i]
    U = torch.rand(n, 2, generator=g)
    ignore_eos = torch.arange(n) % 3 != 0
    return logp, hist, U, ignore_eos


def mel_case(B=2, n=24000, seed=5):
    g = _g(6000 + seed)
    return torch.rand(B, n, generator=g) * 1.6 - 0.8


def z10_utterance(i, n_text=None):
    """canonical ~10 s zero-shot utterance; utterance i of a batch uses seed 1986+i.
    Ragged variant: n_text in {40..60} when n_text is None and i >= 0 (config #3)."""
    g = _g(1986 + i)
    if n_text is None:
        n_text = 50
    text = torch.randint(0, 151643, (1, n_text), dtype=torch.int32, generator=g)
    ptext = torch.randint(0, 151643, (1, 12), dtype=torch.int32, generator=g)
    ptok = torch.randint(0, 6561, (1, 75), dtype=torch.int32, generator=g)
    pfeat = torch.rand(1, 150, 80, generator=g) * 13.5 - 11.5
    emb = torch.randn(1, 192, generator=g)
    return dict(text=text, prompt_text=ptext, llm_prompt_speech_token=ptok, flow_prompt_speech_token=ptok,
                prompt_speech_feat=pfeat, llm_embedding=emb, flow_embedding=emb)
