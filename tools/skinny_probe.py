"""Timeline probe of the LM decode weight-streaming GEMM (clock64 stamps of CTA 0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_b200 import cvk
c = cvk.Context(0, "bf16", 2.0)
c.set_option("debug_timeline", 1)
g = torch.Generator().manual_seed(0)
for (N, K) in ((1152, 896), (9728, 896), (896, 4864), (6564, 896)):
    x = torch.randn(32, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        out, ms, tl = c.linear_small(x, w, None, iters=20, timeline=True)
    ref = x.bfloat16().float() @ w.bfloat16().float().t()
    err = (out.cpu() - ref).abs().max().item()
    t0 = tl[0]
    issue = [t - t0 for t in tl[8:40] if t]
    full = [t - t0 for t in tl[40:72] if t]
    print(f"N={N} K={K}: {ms*1000:.1f} us/iter  ({N*K*2/ms/1e6:.0f} GB/s weights)  maxerr {err:.3g}")
    print("   tma issue  :", issue[:16])
    print("   full ready :", full[:16])
    print("   acc ready  :", tl[1] - t0, " end:", tl[2] - t0)
    import numpy as np
    gg = np.array(tl[128:128 + 800]).reshape(-1, 4)
    gg = gg[gg[:, 0] > 0]
    t00 = gg[:, 0].min()
    print(f"   CTAs {len(gg)}: entry spread {gg[:,0].max()-t00} ns; setup {np.median(gg[:,1]-gg[:,0]):.0f} ns; acc-ready after entry med {np.median(gg[:,2]-gg[:,0]):.0f} max {(gg[:,2]-gg[:,0]).max()} ns; "
          f"end after entry med {np.median(gg[:,3]-gg[:,0]):.0f} max {(gg[:,3]-gg[:,0]).max()} ns; kernel span {gg[:,3].max()-t00} ns")
    print("   epilogue (thread 64, cycles after acc ready): ld0", tl[3] - tl[1], "ld1", tl[4] - tl[1], "stores done", tl[5] - tl[1], "after final sync", tl[6] - tl[1])
