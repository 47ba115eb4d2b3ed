"""Streaming flow stage, per chunk: the reference's prefix recompute (flow.inference(streaming=True, finalize=False) on the growing
prefix, cli/model.py:346-363) against the cached session (cvk_flow_stream_*).  Device time per chunk, one utterance, full-size
random-init CosyVoice2 flow, bf16.  Not the headline bench: a measurement for DESIGN.md / profiles."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cosyvoice_b200 import synth
from cosyvoice_b200.model import B200CosyVoice2Model

ap = argparse.ArgumentParser()
ap.add_argument("--prompt", type=int, default=75, help="prompt speech tokens (3 s)")
ap.add_argument("--tokens", type=int, default=250, help="generated speech tokens (10 s)")
ap.add_argument("--hop", type=int, default=25)
ap.add_argument("--scale", type=int, default=2, help="hop growth factor per chunk (reference: stream_scale_factor), capped at 4 hops")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sds = synth.cosyvoice2_state_dicts(dev, 1986, 1, (6, 4, 12, 4))
m = B200CosyVoice2Model(precision="bf16", device=0, workspace_gb=8.0)
m.load_state_dicts(*sds)
c = m.ctx
g = torch.Generator().manual_seed(5)
P = a.prompt
toks = torch.randint(0, 6561, (P + a.tokens + 3,), generator=g, dtype=torch.int32).to(dev)
pfeat = (torch.rand(2 * P, 80, generator=g) * 13.5 - 11.5).to(dev)
emb = torch.randn(1, 192, generator=g).to(dev)
# chunk schedule of CosyVoice2Model.tts (first hop padded to the 25-token grid, then hop *= scale up to 4 hops)
ends, n, hop = [], P, a.hop
pad = -P % a.hop
while n + (hop + (pad if not ends else 0)) <= P + a.tokens:
    n += hop + (pad if not ends else 0)
    ends.append(n)
    hop = min(4 * a.hop, hop * a.scale)
fs = c.flow_stream(2 * (P + a.tokens) + 64, 10)
print(f"prompt {P} tokens, {a.tokens} generated; chunk ends (tokens) {ends}; session caches {c.flow_stream_bytes(fs) / 2**30:.2f} GiB")


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1), out


res = {"prefix_ms": [], "cached_ms": [], "max_abs_diff": 0.0}
for rep in range(a.reps + 1):
    c.flow_stream_begin(fs, pfeat, emb)
    done, pm, cm = 0, [], []
    for n in ends:
        t_p, (ref, _) = timed(lambda: c.flow_inference(toks[:n + 3], [n + 3], pfeat, [2 * P], emb, streaming=True, finalize=False))
        t_c, new = timed(lambda: c.flow_stream_chunk(fs, toks[:n + 3]))
        want = ref[max(done - 2 * P, 0):]
        res["max_abs_diff"] = max(res["max_abs_diff"], (new - want).abs().max().item())
        done = 2 * n
        pm.append(t_p)
        cm.append(t_c)
    if rep:                      # first repetition = warm-up
        res["prefix_ms"].append(pm)
        res["cached_ms"].append(cm)
mean = lambda rows: [round(sum(r[i] for r in rows) / len(rows), 2) for i in range(len(rows[0]))]
pm, cm = mean(res["prefix_ms"]), mean(res["cached_ms"])
print("per chunk, ms: prefix recompute", pm, " cached session", cm)
print(json.dumps({"chunk_end_tokens": ends, "prefix_recompute_ms": pm, "cached_session_ms": cm, "sum_prefix_ms": round(sum(pm), 1),
                  "sum_cached_ms": round(sum(cm), 1), "speedup": round(sum(pm) / sum(cm), 2), "max_abs_diff": res["max_abs_diff"]}))
c.flow_stream_destroy(fs)
