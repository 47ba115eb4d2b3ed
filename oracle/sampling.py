"""Oracle (test infrastructure): repetition-aware sampling with EXPLICIT uniform draws.

Follows cosyvoice/utils/common.py:138-167 (ras_sampling / nucleus_sampling / random_sampling) and
cosyvoice/llm/llm.py:150-160 (sampling_ids), bound with top_p 0.8, top_k 25, win_size 10, tau_r 0.1
(cosyvoice2.yaml:32-36).

The reference draws with ``Tensor.multinomial`` from the global torch RNG.  To make "bit-exact token ids given
identical logits and identical uniforms" a testable statement, multinomial(1) is DEFINED here as an inverse-CDF
lookup on a supplied uniform u in [0,1) (``draw_index``); tests monkey-patch ``torch.Tensor.multinomial`` with the
same rule when running the reference, so reference, oracle and CUDA kernel consume identical draws.  All arithmetic
is float32 in a fixed order so that the CUDA kernel can reproduce it bit for bit.
"""
import numpy as np

TOP_P, TOP_K, WIN_SIZE, TAU_R = 0.8, 25, 10, 0.1
SEG = 256          # the CUDA kernel uses 256 threads, each owning a contiguous segment


def draw_index(w, u):
    """Inverse-CDF draw from unnormalised non-negative float32 weights ``w`` with one uniform ``u``.

    Hierarchical float32 CDF (matches the kernel): the vector is cut into SEG contiguous segments of
    ceil(n/SEG) entries; segment sums are accumulated left to right, then segment prefix sums left to right;
    threshold = u * total; the chosen entry is the first whose running sum exceeds the threshold."""
    w = np.asarray(w, dtype=np.float32)
    n = w.shape[0]
    per = (n + SEG - 1) // SEG
    seg_sum = np.zeros(SEG, dtype=np.float32)
    for s in range(SEG):
        acc = np.float32(0)
        for i in range(s * per, min((s + 1) * per, n)):
            acc = np.float32(acc + w[i])
        seg_sum[s] = acc
    prefix = np.zeros(SEG + 1, dtype=np.float32)
    for s in range(SEG):
        prefix[s + 1] = np.float32(prefix[s] + seg_sum[s])
    thr = np.float32(np.float32(u) * prefix[SEG])
    last = -1
    for s in range(SEG):
        if seg_sum[s] <= 0:
            continue
        if prefix[s + 1] > thr:
            acc = prefix[s]
            for i in range(s * per, min((s + 1) * per, n)):
                if w[i] > 0:
                    last = i
                acc = np.float32(acc + w[i])
                if acc > thr and w[i] > 0:
                    return i
        for i in range(s * per, min((s + 1) * per, n)):
            if w[i] > 0:
                last = i
    return last


def hsum(w):
    """float32 sum in the kernel's order: SEG contiguous segments summed left to right, then the SEG partial
    sums left to right."""
    w = np.asarray(w, dtype=np.float32)
    n = w.shape[0]
    per = (n + SEG - 1) // SEG
    tot = np.float32(0)
    for s in range(SEG):
        acc = np.float32(0)
        for i in range(s * per, min((s + 1) * per, n)):
            acc = np.float32(acc + w[i])
        tot = np.float32(tot + acc)
    return tot


def softmax_f32(x):
    """float32 softmax with -inf support: exp(x - max) / hsum."""
    x = np.asarray(x, dtype=np.float32)
    m = np.max(x)
    e = np.exp((x - m).astype(np.float32)).astype(np.float32)
    e[np.isneginf(x)] = 0
    return (e / hsum(e)).astype(np.float32)


def log_softmax_f32(x):
    """float32 log-softmax the way the decode kernel forms it: (x - max) - log(hsum(exp(x - max)))."""
    x = np.asarray(x, dtype=np.float32)
    m = np.max(x)
    d = (x - m).astype(np.float32)
    return (d - np.log(hsum(np.exp(d).astype(np.float32))).astype(np.float32)).astype(np.float32)


def nucleus_select(prob):
    """common.py:147-162: stable descending sort, keep entries while cum < top_p and n < top_k (the entry that
    crosses top_p is kept).  Returns (indices, probs) of the kept prefix."""
    order = np.argsort(-prob.astype(np.float64), kind="stable")
    cum = np.float32(0)
    idx, pr = [], []
    for i in order:
        if cum < np.float32(TOP_P) and len(pr) < TOP_K:
            cum = np.float32(cum + prob[i])
            pr.append(prob[i])
            idx.append(int(i))
        else:
            break
    return np.array(idx, dtype=np.int64), np.array(pr, dtype=np.float32)


def ras_sample(scores, decoded_tokens, u1, u2, ignore_eos, eos_index=6561):
    """llm.py:150-160 + common.py:138-144.  scores: log-probs [V]; returns the sampled token id (int)."""
    scores = np.array(scores, dtype=np.float32)
    if ignore_eos:
        scores[eos_index] = -np.inf
    prob = softmax_f32(scores)
    idx, pr = nucleus_select(prob)
    top = int(idx[draw_index(pr, u1)])
    rep = sum(1 for t in decoded_tokens[-WIN_SIZE:] if t == top)
    if rep >= WIN_SIZE * TAU_R:
        scores[top] = -np.inf
        top = int(draw_index(softmax_f32(scores), u2))
    return top
