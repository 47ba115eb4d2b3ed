"""bf16 / 24-layer sanity run of the text-streaming LM (timing of the host-driven path; not a bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cases, lm
from cosyvoice_b200.model import B200CosyVoice2Model
chunks, ptext, ptok, U = cases.bistream_case()
m = B200CosyVoice2Model(precision="bf16", device=0, workspace_gb=4.0)
m.ctx.load_state_dict("llm", lm.bistream_state_dict(24), [24])
torch.cuda.synchronize()
t0 = time.perf_counter()
ids = list(m.lm_generate_bistream(iter(chunks), ptext, ptok, uniforms=U))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"bf16 24-layer bistream: {len(ids)} ids in {dt * 1e3:.0f} ms ({dt / max(len(ids), 1) * 1e3:.2f} ms per id, {len(ids) * 0.04 / dt:.1f} audio-s/s single stream), all < 6561: {all(i < 6561 for i in ids)}")
