"""CPU: the index algebra of two weight repacks used by the CosyVoice3 causal vocoder (cosyvoice_b200/csrc/hift.cu), restated
in Python exactly as the CUDA kernels compute it and checked against the torch ops of the reference.  The conv-GEMM contract is
out[r, n] = bias[n] + sum_j sum_k A[r + shift0 + j, k] * W[n][j][k] on zero-padded rows."""
import torch
import torch.nn.functional as F


def floor_div(a, b):
    return a // b            # Python floors like the kernel's floor_div


def conv_gemm(A, W, shift0):
    """A [R,K], W [N,taps,K] -> [R,N] with rows outside [0,R) reading zeros"""
    R, K = A.shape
    N, taps, _ = W.shape
    out = torch.zeros(R, N, dtype=A.dtype)
    for j in range(taps):
        for r in range(R):
            rr = r + shift0 + j
            if 0 <= rr < R:
                out[r] += W[:, j, :] @ A[rr]
    return out


def test_nearest_upsample_causal_conv_as_polyphase():
    """upsample_causal_poly_kernel: nn.Upsample(nearest, u) + F.pad(k-1, 0) + Conv1d(k) (convolution.py:224-258) == a 3-tap conv
    with shift0 = -2 producing u*Cout columns whose [R, u*Cout] matrix is the [u*R, Cout] output"""
    g = torch.Generator().manual_seed(0)
    for u, k in ((8, 16), (5, 11), (3, 7)):
        Cin, Cout, T = 6, 4, 9
        w = torch.randn(Cout, Cin, k, generator=g, dtype=torch.float64)
        x = torch.randn(1, Cin, T, generator=g, dtype=torch.float64)
        ref = F.conv1d(F.pad(x.repeat_interleave(u, dim=2), (k - 1, 0)), w)[0].t()          # [u*T, Cout]
        W = torch.zeros(u * Cout, 3, Cin, dtype=torch.float64)
        for n in range(u * Cout):
            p, co = n // Cout, n % Cout
            for jt in range(3):
                for j in range(k):
                    if floor_div(p - (k - 1) + j, u) == jt - 2:
                        W[n, jt] += w[co, :, j]
        out = conv_gemm(x[0].t().contiguous(), W, -2).reshape(T * u, Cout)                  # [T, u*Cout] viewed as [u*T, Cout]
        assert torch.allclose(out, ref, atol=1e-12), (u, k, (out - ref).abs().max())


def test_causal_strided_source_down_on_the_stft_view():
    """strided_view_fill_kernel with off = stride: CausalConv1dDownSample(k = 2*stride, stride) (convolution.py:190-221) over the
    STFT matrix [F,18] == a 3-tap conv (shift0 = -1) over its [F'/s, s*24] view.  Level-3 row of frame f is f - 1 relative to
    s * (first output row) (the vocoder's level-3 geometry has one extra front row)."""
    g = torch.Generator().manual_seed(1)
    LD = 24
    for s in (15, 3):
        k, N, Tout = 2 * s, 5, 7
        nfr = s * Tout + 1                                                                 # STFT frames 0 .. s*Tout
        w = torch.randn(N, 18, k, generator=g, dtype=torch.float64)
        stft = torch.randn(nfr, 18, generator=g, dtype=torch.float64)
        ref = F.conv1d(F.pad(stft.t()[None], (s - 1, 0)), w, stride=s)[0].t()              # [Tout, N]
        assert ref.shape[0] == Tout
        # level-3 storage: frame f at row (s*start + f - 1) with start = 1 output row of margin in front
        start = 1
        rows3 = s * (start + Tout + 1)
        M = torch.zeros(rows3, LD, dtype=torch.float64)
        M[s * start - 1: s * start - 1 + nfr, :18] = stft
        view = M.reshape(rows3 // s, s * LD)
        Kv = s * LD
        W = torch.zeros(N, 3, Kv, dtype=torch.float64)
        off = s
        for n in range(N):
            for c in range(18):
                for j in range(k):
                    jp = j - off
                    dq = floor_div(jp, s)
                    pp = jp - dq * s
                    W[n, dq + 1, pp * LD + c] = w[n, c, j]
        out = conv_gemm(view, W, -1)[start:start + Tout]
        assert torch.allclose(out, ref, atol=1e-12), (s, (out - ref).abs().max())


def conv_gemm_fast(A, W, shift0, dil=1, bias=None):
    """vectorised conv_gemm: A [R,K], W [N,taps,K] -> [R,N]; rows outside [0,R) read zeros"""
    R = A.shape[0]
    out = torch.zeros(R, W.shape[0], dtype=A.dtype)
    for j in range(W.shape[1]):
        off = shift0 + j * dil
        lo, hi = max(0, -off), min(R, R - off)
        if hi > lo:
            out[lo:hi] += A[lo + off:hi + off] @ W[:, j, :].t()
    return out if bias is None else out + bias


def _causal_body(finalize):
    """The CosyVoice3 vocoder body exactly as cosyvoice_b200/csrc/hift.cu wires it (conv_pre looking right, three polyphase
    nearest-up-sampling conv-GEMMs whose [R, u*C] output is reinterpreted as [u*R, C], the reflected front row at the last level,
    strided source_downs on the [R/s, s*24] STFT view with the front-row offset, left-shifted dilated ResBlock convs, causal
    conv_post), in float64 on one short utterance, against the oracle's decode (pinned to the reference CausalHiFTGenerator)."""
    from oracle import cases, hift_causal as hc, hift as H, weights
    sd32 = weights.synth_state_dict(hc.param_shapes(), 1986, hc.SYNTH_GAINS)
    sd = {k: v.double() for k, v in sd32.items()}
    g = torch.Generator().manual_seed(3)
    # streaming call: the vocoder receives the mel without the 3 f0 look-ahead frames (generator.py:725); conv_pre then uses 4 more
    # frames as look-ahead, so the body runs on n_body = T - 4 rows while the source (and its STFT, before the cut) covers T rows
    T = 5 if finalize else 10
    mel32 = torch.randn(1, 80, T, generator=g) * 2 - 5
    src32 = torch.tanh(torch.randn(1, 1, T * 480, generator=g))
    mel, src = mel32.double(), src32.double()
    n_body = T if finalize else T - 4
    ref = hc.decode(sd32, mel32, src32, finalize, return_pre_istft=True)[0].t().double()    # [120 n_body + 1, 18], oracle in fp32

    def wn(prefix):                                                                          # effective weight [out, in, k]
        return H._w(sd, prefix)

    def as_gemm(w):                                                                          # [out, in, k] -> [out, k, in]
        return w.permute(0, 2, 1).contiguous()
    snake = lambda x, a: x + (1.0 / (a[None, :] + 1e-9)) * torch.sin(x * a[None, :]) ** 2
    lrelu = torch.nn.functional.leaky_relu
    # STFT of the source, frames x 18, stored at level-3 rows f (front row included)
    re, im = H.stft16(src.squeeze(1).float())
    stft = torch.cat([re, im], 1)[0].t().double()[:120 * n_body + 1]                          # frames kept after the cut (generator.py:681-682)
    # conv_pre over ALL rows (it looks 4 rows to the right, into the look-ahead rows of the same matrix)
    x = lrelu(conv_gemm_fast(mel[0].t(), as_gemm(wn("conv_pre")), 0, bias=sd["conv_pre.bias"]), 0.1)
    x = x[:n_body]                       # later rows are masked to zero on the device and never read by the causal layers below
    ups, upk, ch = (8, 5, 3), (16, 11, 7), (512, 256, 128, 64)
    strides = (15, 3, 1)
    for i in range(3):
        u, k, Cin, Cout = ups[i], upk[i], ch[i], ch[i + 1]
        w = wn(f"ups.{i}")
        W = torch.zeros(u * Cout, 3, Cin, dtype=torch.float64)
        for p in range(u):
            for jt in range(3):
                for j in range(k):
                    if (p - (k - 1) + j) // u == jt - 2:
                        W[p * Cout:(p + 1) * Cout, jt] += w[:, :, j]
        xu = conv_gemm_fast(x, W, -2, bias=sd[f"ups.{i}.bias"].repeat(u)).reshape(x.shape[0] * u, Cout)
        if i == 2:
            xu = torch.cat([xu[1:2], xu], 0)                                                 # reflect_front_kernel
        R = xu.shape[0]
        # source branch on the strided view
        s = strides[i]
        ws = sd[f"source_downs.{i}.weight"]
        if s == 1:
            si = stft @ ws[:, :, 0].t() + sd[f"source_downs.{i}.bias"]
        else:
            LD, start = 24, 1
            rows3 = s * (start + R + 1)
            M = torch.zeros(rows3, LD, dtype=torch.float64)
            M[s * start - 1: s * start - 1 + stft.shape[0], :18] = stft
            Wv = torch.zeros(ws.shape[0], 3, s * LD, dtype=torch.float64)
            for c in range(18):
                for j in range(ws.shape[2]):
                    jp = j - s
                    dq = jp // s
                    Wv[:, dq + 1, (jp - dq * s) * LD + c] = ws[:, c, j]
            si = conv_gemm_fast(M.reshape(rows3 // s, s * LD), Wv, -1, bias=sd[f"source_downs.{i}.bias"])[start:start + R]

        def resblock(prefix, x, k):
            for d_i, d in enumerate((1, 3, 5)):
                xt = snake(x, sd[f"{prefix}.activations1.{d_i}.alpha"])
                xt = conv_gemm_fast(xt, as_gemm(wn(f"{prefix}.convs1.{d_i}")), -(k - 1) * d, d, sd[f"{prefix}.convs1.{d_i}.bias"])
                xt = snake(xt, sd[f"{prefix}.activations2.{d_i}.alpha"])
                xt = conv_gemm_fast(xt, as_gemm(wn(f"{prefix}.convs2.{d_i}")), -(k - 1), 1, sd[f"{prefix}.convs2.{d_i}.bias"])
                x = xt + x
            return x
        xu = xu + resblock(f"source_resblocks.{i}", si, (7, 7, 11)[i])
        xs = sum(resblock(f"resblocks.{i * 3 + j}", xu, kk) for j, kk in enumerate((3, 7, 11)))
        x = lrelu(xs / 3, 0.1 if i < 2 else 0.01)
    out = conv_gemm_fast(x, as_gemm(wn("conv_post")), -6, bias=sd["conv_post.bias"])[:ref.shape[0]]
    assert (out - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item()), ((out - ref).abs().max(), ref.abs().max())


def test_causal_vocoder_body_as_conv_gemms():
    _causal_body(True)


def test_causal_vocoder_body_streaming_call():
    """finalize=False: the body rows are a prefix of the mel rows (shared row starts), conv_pre reads its 4 look-ahead rows from the
    same matrix, the STFT of the full source is cut to 120 n_body + 1 frames - everything downstream is causal, so the masked tail
    rows never reach a kept output"""
    _causal_body(False)
