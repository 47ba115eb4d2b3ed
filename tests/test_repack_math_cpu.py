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
