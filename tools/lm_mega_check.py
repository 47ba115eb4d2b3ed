"""Persistent LM decode kernel (llm_mega.cu): numeric check against the per-op fused chain + step time + phase timeline.  Not a bench."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cosyvoice_b200 import synth
from cosyvoice_b200.model import B200CosyVoice2Model

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--coop", type=int, default=1)
ap.add_argument("--timeline", type=int, default=1)
ap.add_argument("--a", default="lm_mega=1", help="options of variant A (comma-separated key=value)")
ap.add_argument("--b", default="lm_mega=0", help="options of variant B (the reference of the comparison)")
a = ap.parse_args()


def parse_opts(txt):
    return {k: int(v) for k, v in (kv.split("=") for kv in txt.split(",") if kv)}


OPT_A, OPT_B = parse_opts(a.a), parse_opts(a.b)
dev = torch.device("cuda", 0)
sds = synth.cosyvoice2_state_dicts(dev, 1986, a.layers, (2, 1, 2, 2))
m = B200CosyVoice2Model(precision="bf16", device=0, workspace_gb=8.0)
m.load_state_dicts(*sds)
c = m.ctx
c.set_option("mega_coop", a.coop)
inputs = synth.batch32_zero_shot(a.batch)
B = a.batch
texts = [i["text"] for i in inputs]
ptexts = [i["prompt_text"] for i in inputs]
ptoks = [i["llm_prompt_speech_token"] for i in inputs]
tl = [int(t.shape[1] + p.shape[1]) for t, p in zip(texts, ptexts)]
sl = [int(s.shape[1]) for s in ptoks]
tt = torch.cat([torch.cat([p, t], 1).reshape(-1) for t, p in zip(texts, ptexts)])
ss = torch.cat([s.reshape(-1) for s in ptoks])
g = torch.Generator().manual_seed(3)
U = torch.rand(a.steps + 8, B, 2, generator=g)
U[:, :, 0] = 1e-4                                       # nucleus draw lands on the most probable id
Ud = U.to(dev)
big = torch.full((B,), 100000, dtype=torch.int32, device=dev)


def run(opts, nsteps, timeline=False):
    for k, v in opts.items():
        c.set_option(k, v)
    sess = c.lm_session(B, max(tl) + max(sl) + a.steps + 32)
    ids = torch.zeros(B, a.steps + 8, dtype=torch.int32, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    done = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        c.lm_prefill(sess, tt, tl, ss, sl)
        c.lm_decode(sess, 2, Ud, big, big, ids, cnt, done)
        lg2 = c.lm_last_logits(sess, B).clone()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c.lm_decode(sess, nsteps, Ud, big, big, ids, cnt, done, want_live=False)
        e1.record()
        st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / max(nsteps, 1)
        if timeline:
            c.set_option("chain_timeline", 1)
            c.lm_decode(sess, 1, Ud, big, big, ids, cnt, done)       # graph is re-captured? no: same args -> replay; stamps need tl in params
            st.synchronize()
        lg = c.lm_last_logits(sess, B).clone()
        out = (lg2.cpu(), lg.cpu(), ids.cpu().clone(), us)
    torch.cuda.synchronize()
    c.lm_session_destroy(sess)
    return out


l2a, la, ia, usa = run(OPT_A, a.steps)
l2b, lb, ib, usb = run(OPT_B, a.steps)
fin = torch.isfinite(l2a) & torch.isfinite(l2b)
print(f"B={B} layers={a.layers}: step time A {OPT_A} {usa:.1f} us, B {OPT_B} {usb:.1f} us")
print(f"logits after 2 steps: max |mega - chain| = {(l2a - l2b)[fin].abs().max().item():.4g} (|logit| up to {l2b[fin].abs().max().item():.3g}), finite {fin.float().mean().item():.4f}")
same = (ia == ib)
first_div = [int((~same[b]).nonzero()[0]) if (~same[b]).any() else -1 for b in range(B)]
print("first diverging step per row (-1 = never):", first_div)
fin = torch.isfinite(la) & torch.isfinite(lb)
print(f"logits after {a.steps + 2} steps: max |mega - chain| = {(la - lb)[fin].abs().max().item():.4g}")
if a.timeline:
    # timeline: needs the option set BEFORE the graph is captured
    for k, v in OPT_A.items():
        c.set_option(k, v)
    c.set_option("chain_timeline", 1)
    sess = c.lm_session(B, max(tl) + max(sl) + 64)
    ids = torch.zeros(B, 64, dtype=torch.int32, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    done = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        c.lm_prefill(sess, tt, tl, ss, sl)
        c.lm_decode(sess, 24, Ud, big, big, ids, cnt, done)
        st.synchronize()
    t = c.debug_read(4096)
    mg = [x for x in t[2048:2048 + 7 * a.layers + 2]]
    names = ["qkv", "attn", "o", "red1", "gate_up", "down", "red2"]
    if mg[0]:
        d = [(mg[i + 1] - mg[i]) / 1e3 for i in range(len(mg) - 1) if mg[i + 1]]
        print(f"mega kernel: {len(d)} phases, total {(mg[len(d)] - mg[0]) / 1e3:.1f} us")
        for k, n in enumerate(names):
            v = d[k::7]
            print(f"  {n:8s} mean {sum(v) / len(v):6.2f} us  (layer 0: {v[0]:.2f}, layer 1: {v[1] if len(v) > 1 else 0:.2f}, last: {v[-1]:.2f})")
    if not OPT_A.get("lm_mega", 0):
        # per-op chain: a block of 4 slots per kernel, 3 stamps (entry, past griddepcontrol.wait, end) per kernel in launch order
        st3 = [t[i * 4:i * 4 + 3] for i in range(255) if t[i * 4]]
        per_layer = (len(st3) - 3) // a.layers if a.layers else 0
        print(f"chain: {len(st3)} stamped kernels, {per_layer} per layer")
        if per_layer:
            for k in range(per_layer):
                rows = [st3[3 + l * per_layer + k] for l in range(1, a.layers)]
                nxt = [st3[3 + l * per_layer + k + 1] if 3 + l * per_layer + k + 1 < len(st3) else None for l in range(1, a.layers)]
                span = [(r[2] - r[1]) / 1e3 for r in rows]
                period = [(n[1] - r[1]) / 1e3 for r, n in zip(rows, nxt) if n]
                print(f"  kernel {k}: waited->end {sum(span) / len(span):5.2f} us, waited->next waited {sum(period) / max(len(period), 1):5.2f} us")
            l0, l1 = st3[3 + per_layer][1], st3[3 + (a.layers - 1) * per_layer][1]
            print(f"  layer period {(l1 - l0) / 1e3 / (a.layers - 2):.2f} us")
    ch = t[:16]
    print("chain stamps (head, head_finish, sampler) entry/waited/end us rel:", [round((x - ch[0]) / 1e3, 2) if x else 0 for x in ch[:12]])
    if mg[0] and ch[0]:
        print(f"sampler past pdl_wait -> mega first stamp: {(mg[0] - ch[9]) / 1e3:.2f} us; step (head waited -> mega end): {(mg[7 * a.layers] - ch[1]) / 1e3:.1f} us")

    for name, off in (("CTA 0", 2048 + 512), ("CTA 100", 2048 + 768)):
        f = [x for x in t[off:off + 250] if x]
        if f:
            print(f"fine stamps layer 1, {name} ({len(f)}): deltas us:", [round((f[i + 1] - f[i]) / 1e3, 2) for i in range(len(f) - 1)])
    print("sampler end:", round((ch[10] - ch[0]) / 1e3, 2) if ch[10] else 0)

    for name, off, n in (("sampler stages (CTA 0)", 1024, 16), ("MMA thread gate_up L1 (CTA 0)", 2048 + 1280, 40), ("epilogue thread gate_up L1 (CTA 0)", 2048 + 1400, 8)):
        f = [x for x in t[off:off + n] if x]
        if f:
            print(f"{name} ({len(f)} stamps): deltas us:", [round((f[i + 1] - f[i]) / 1e3, 2) for i in range(len(f) - 1)])
