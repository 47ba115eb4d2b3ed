"""Times the tcgen05 conv-GEMM on the estimator / HiFT shapes under different tile / epilogue options."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosyvoice_b200 import cvk
c = cvk.Context(0, "bf16", 12.0)
c.set_option("op_iters", 10)
c.set_option("debug_timeline", 1)
import numpy as np
g = torch.Generator().manual_seed(0)
SHAPES = [  # name, rows, K, N, taps, dil, act
    ("est ff1 gelu", 40064, 256, 1024, 1, 1, "gelu"), ("est ff2", 40064, 1024, 256, 1, 1, "none"), ("est qkv", 40064, 256, 1536, 1, 1, "none"),
    ("est conv k3", 40064, 256, 256, 3, 1, "none"), ("est to_out", 40064, 512, 256, 1, 1, "none"),
    ("hift L1 k7", 130000, 256, 256, 7, 3, "snake"), ("hift L3 k11", 1900000, 64, 64, 11, 5, "none"), ("hift L2 k3", 650000, 128, 128, 3, 1, "none")]
for name, rows, K, N, taps, dil, act in SHAPES:
    x = torch.randn(rows, K, generator=g)
    w = torch.randn(N, K, taps, generator=g) / (K * taps) ** 0.5
    b = torch.randn(N, generator=g)
    a = act if act != "snake" else "silu"
    res = []
    for bn256, epi in itertools.product((0,), (0, 2)):
        c.set_option("tc_bn256", bn256); c.set_option("tc_epi", epi)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            c.conv1d(x, [rows], w, b, dil=dil, shift0=-(taps - 1) * dil // 2, act=a)
        ms = c.last_op_ms()
        res.append(f"bn256={bn256} epi={epi}: {ms*1000:7.1f} us {2*rows*N*K*taps/ms/1e9:6.0f} TF")
        if epi == 2:
            t = np.array(c.debug_read(960)).reshape(-1, 8)
            t = t[t[:, 0] > 0]
            e0 = t[:, 0]
            f = lambda k: f"{np.median(t[:, k] - e0):.0f}"
            res.append(f"[ns after CTA entry, median of {len(t)} CTAs: setup {f(1)} first-operands {f(6)} acc-ready {f(2)} math+stage done {f(3)} barrier {f(4)} tma-store done {f(5)}; entry spread {e0.max()-e0.min()}]")
    print(f"{name:14s} M={rows} K={K} N={N} taps={taps}: " + " | ".join(res))
