"""Oracle (test infrastructure): HiFT vocoder (CosyVoice2 config) restated in plain torch fp32 on CPU.

Follows cosyvoice/hifigan/generator.py (HiFTGenerator.inference :557-569, decode :507-539,
_stft/_istft :491-505, SourceModuleHnNSF.forward :358-375, SineGen2 :233-317, ResBlock.forward :110-117),
cosyvoice/hifigan/f0_predictor.py:56-59 and cosyvoice/transformer/activation.py:73-84 (Snake), with the
hyper-parameters of examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:89-111.

All randomness is an explicit input (the reference draws ``rand_ini`` and Gaussian noise from the global
torch RNG inside SineGen2, generator.py:243-247,306-309), so that reference, oracle and CUDA consume the
same draws.  Pinned against the reference module by oracle/make_golden.py (tests/golden/hift_*.npz).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from .weights import weight_norm_effective

UPS_RATES = [8, 5, 3]
UPS_KERNELS = [16, 11, 7]
RB_KERNELS = [3, 7, 11]
RB_DILS = [1, 3, 5]
SRC_RB_KERNELS = [7, 7, 11]
BASE_CH = 512
N_FFT, HOP = 16, 4
NB_HARM = 8
SR = 24000
UPSCALE = 480           # prod(UPS_RATES) * HOP  (generator.py:415)
LRELU = 0.1
AUDIO_LIMIT = 0.99
SINE_AMP, NOISE_STD, VOICED_THR = 0.1, 0.003, 10.0


def _wn(shapes, prefix, wshape):
    shapes[prefix + ".bias"] = (wshape[0],) if "ups." not in prefix else (wshape[1],)
    shapes[prefix + ".parametrizations.weight.original0"] = (wshape[0], 1, 1)
    shapes[prefix + ".parametrizations.weight.original1"] = tuple(wshape)


def _resblock_shapes(shapes, prefix, ch, k):
    for i in range(3):
        _wn(shapes, f"{prefix}.convs1.{i}", (ch, ch, k))
    for i in range(3):
        _wn(shapes, f"{prefix}.convs2.{i}", (ch, ch, k))
    for i in range(3):
        shapes[f"{prefix}.activations1.{i}.alpha"] = (ch,)
    for i in range(3):
        shapes[f"{prefix}.activations2.{i}.alpha"] = (ch,)


def param_shapes():
    """state_dict keys/shapes of the reference HiFTGenerator, in module order."""
    s = OrderedDict()
    s["m_source.l_linear.weight"] = (1, NB_HARM + 1)
    s["m_source.l_linear.bias"] = (1,)
    _wn(s, "conv_pre", (BASE_CH, 80, 7))
    for i, (u, k) in enumerate(zip(UPS_RATES, UPS_KERNELS)):
        _wn(s, f"ups.{i}", (BASE_CH // 2 ** i, BASE_CH // 2 ** (i + 1), k))
    down = [(30, 15), (6, 3), (1, 1)]
    for i in range(3):
        ch = BASE_CH // 2 ** (i + 1)
        s[f"source_downs.{i}.weight"] = (ch, N_FFT + 2, down[i][0])
        s[f"source_downs.{i}.bias"] = (ch,)
    for i in range(3):
        _resblock_shapes(s, f"source_resblocks.{i}", BASE_CH // 2 ** (i + 1), SRC_RB_KERNELS[i])
    for i in range(3):
        for j, k in enumerate(RB_KERNELS):
            _resblock_shapes(s, f"resblocks.{i * 3 + j}", BASE_CH // 2 ** (i + 1), k)
    _wn(s, "conv_post", (N_FFT + 2, BASE_CH // 8, 7))
    cin = 80
    for i in range(5):
        _wn(s, f"f0_predictor.condnet.{2 * i}", (512, cin, 3))
        cin = 512
    s["f0_predictor.classifier.weight"] = (1, 512)
    s["f0_predictor.classifier.bias"] = (1,)
    return s


# scale rules that keep the random-weight vocoder in a non-degenerate regime (|conv_post| ~ 1)
SYNTH_GAINS = {"conv_post.parametrizations.weight.original0": 0.15, "f0_predictor.classifier.weight": 60.0,
               "f0_predictor.classifier.bias": 2000.0}


def _w(sd, prefix):
    return weight_norm_effective(sd[prefix + ".parametrizations.weight.original0"],
                                 sd[prefix + ".parametrizations.weight.original1"])


def snake(x, alpha):
    """activation.py:73-84: x + 1/(a+1e-9) * sin^2(a x), per channel."""
    a = alpha.view(1, -1, 1)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def resblock(sd, prefix, x, k):
    """generator.py:110-117."""
    for i, d in enumerate(RB_DILS):
        xt = snake(x, sd[f"{prefix}.activations1.{i}.alpha"])
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs1.{i}"), sd[f"{prefix}.convs1.{i}.bias"],
                      dilation=d, padding=(k * d - d) // 2)
        xt = snake(xt, sd[f"{prefix}.activations2.{i}.alpha"])
        xt = F.conv1d(xt, _w(sd, f"{prefix}.convs2.{i}"), sd[f"{prefix}.convs2.{i}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def f0_predict(sd, mel):
    """f0_predictor.py:56-59.  mel [B,80,T] -> f0 [B,T] (>= 0)."""
    x = mel
    for i in range(5):
        p = f"f0_predictor.condnet.{2 * i}"
        x = F.elu(F.conv1d(x, _w(sd, p), sd[p + ".bias"], padding=1))
    x = x.transpose(1, 2)
    return torch.abs(F.linear(x, sd["f0_predictor.classifier.weight"], sd["f0_predictor.classifier.bias"]).squeeze(-1))


def sine_source(sd, f0, noise, rand_ini=None):
    """generator.py:560-564 + SourceModuleHnNSF :358-375 + SineGen2 :233-317 (non-causal, eval).

    f0 [B,T]; noise [B, 480T, 9] standard normal draws (the reference's randn_like, :309);
    rand_ini [B,9] initial phase draws (:246-247; provably without effect for upsample_scale 480 because the
    linear 1/480 down-sampling reads samples 480d+239 and 480d+240 only - kept for fidelity).
    returns s [B,1,480T]."""
    B, T = f0.shape
    L = T * UPSCALE
    f0u = f0[:, :, None].repeat_interleave(UPSCALE, dim=1)                 # nn.Upsample(nearest) [B,L,1]
    harm = torch.arange(1, NB_HARM + 2, dtype=torch.float32).view(1, 1, -1)
    fn = f0u * harm
    rad = (fn / SR) % 1
    if rand_ini is not None:
        rad[:, 0, :] = rad[:, 0, :] + rand_ini
    rad = F.interpolate(rad.transpose(1, 2), scale_factor=1 / UPSCALE, mode="linear").transpose(1, 2)
    phase = torch.cumsum(rad, dim=1) * 2 * np.pi
    phase = F.interpolate(phase.transpose(1, 2) * UPSCALE, scale_factor=UPSCALE, mode="linear").transpose(1, 2)
    sines = torch.sin(phase) * SINE_AMP
    uv = (f0u > VOICED_THR).float()
    noise_amp = uv * NOISE_STD + (1 - uv) * SINE_AMP / 3
    sine_waves = sines * uv + noise_amp * noise
    merged = torch.tanh(F.linear(sine_waves, sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))
    return merged.transpose(1, 2)


def hann16():
    """scipy.signal.get_window('hann', 16, fftbins=True) == periodic hann (generator.py:474)."""
    n = torch.arange(N_FFT, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * n / N_FFT)).float()


def stft16(s):
    """generator.py:491-497 as an explicit DFT.  s [B,L] -> (re, im) each [B,9,L/4+1]."""
    w = hann16()
    x = F.pad(s.unsqueeze(1), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(1)
    frames = x.unfold(1, N_FFT, HOP) * w                                   # [B, F, 16]
    n = torch.arange(N_FFT, dtype=torch.float64)
    f = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * f[:, None] * n[None, :] / N_FFT
    re = torch.einsum("bfn,kn->bkf", frames.double(), torch.cos(ang)).float()
    im = torch.einsum("bfn,kn->bkf", frames.double(), -torch.sin(ang)).float()
    return re, im


def istft16(mag, phase):
    """generator.py:499-505 as explicit inverse DFT + windowed overlap-add.  [B,9,F] -> [B,4(F-1)]."""
    mag = torch.clip(mag, max=1e2)
    re = mag * torch.cos(phase)
    im = mag * torch.sin(phase)
    B, K, Fr = re.shape
    w = hann16().double()
    n = torch.arange(N_FFT, dtype=torch.float64)
    k = torch.arange(K, dtype=torch.float64)
    ang = 2 * math.pi * k[:, None] * n[None, :] / N_FFT
    coef = torch.full((K, 1), 2.0, dtype=torch.float64)
    coef[0] = 1.0
    coef[-1] = 1.0
    cr = coef * torch.cos(ang) / N_FFT                                     # irfft: imag of DC/Nyquist ignored
    ci = -coef * torch.sin(ang) / N_FFT
    ci[0] = 0
    ci[-1] = 0
    frames = torch.einsum("bkf,kn->bfn", re.double(), cr) + torch.einsum("bkf,kn->bfn", im.double(), ci)
    frames = frames * w
    total = HOP * (Fr - 1) + N_FFT
    y = torch.zeros(B, total, dtype=torch.float64)
    env = torch.zeros(total, dtype=torch.float64)
    for t in range(Fr):
        y[:, t * HOP:t * HOP + N_FFT] += frames[:, t]
        env[t * HOP:t * HOP + N_FFT] += w * w
    y = y[:, N_FFT // 2: total - N_FFT // 2] / env[N_FFT // 2: total - N_FFT // 2]
    return y.float()


def decode(sd, mel, s, return_pre_istft=False):
    """generator.py:507-539.  mel [B,80,T], s [B,1,480T] -> wav [B,480T]."""
    re, im = stft16(s.squeeze(1))
    s_stft = torch.cat([re, im], dim=1)
    x = F.conv1d(mel, _w(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    down = [(15, 7), (3, 1), (1, 0)]
    for i in range(3):
        x = F.leaky_relu(x, LRELU)
        u, k = UPS_RATES[i], UPS_KERNELS[i]
        x = F.conv_transpose1d(x, _w(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i == 2:
            x = F.pad(x, (1, 0), mode="reflect")
        si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"],
                      stride=down[i][0], padding=down[i][1])
        si = resblock(sd, f"source_resblocks.{i}", si, SRC_RB_KERNELS[i])
        x = x + si
        xs = None
        for j, k in enumerate(RB_KERNELS):
            r = resblock(sd, f"resblocks.{i * 3 + j}", x, k)
            xs = r if xs is None else xs + r
        x = xs / 3
    x = F.leaky_relu(x)                                                     # default slope 0.01 (:532)
    x = F.conv1d(x, _w(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    if return_pre_istft:
        return x
    mag = torch.exp(x[:, :N_FFT // 2 + 1])
    phase = torch.sin(x[:, N_FFT // 2 + 1:])
    y = istft16(mag, phase)
    return torch.clamp(y, -AUDIO_LIMIT, AUDIO_LIMIT)


def inference(sd, mel, noise, rand_ini=None, cache_source=None):
    """generator.py:557-569.  Returns (wav [B,480T], source [B,1,480T])."""
    f0 = f0_predict(sd, mel)
    s = sine_source(sd, f0, noise, rand_ini)
    if cache_source is not None and cache_source.shape[2] != 0:
        s[:, :, :cache_source.shape[2]] = cache_source
    return decode(sd, mel, s), s
