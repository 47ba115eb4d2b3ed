"""Stage-by-stage localisation of a tts() mismatch against the oracle (debug aid; uses oracle/, never shipped)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import cases, flow, hift, lm, weights
from oracle.make_golden import stream_noise
from cosyvoice_b200.model import B200CosyVoice2Model

NL, kw = 2, dict(enc_blocks=2, enc_up_blocks=1, num_mid_blocks=2, n_blocks=2)
cfg = flow.FlowCfg(**kw)
sd_l = lm.synth_state_dict(NL)
sd_f = weights.synth_state_dict(flow.param_shapes(cfg), 1986, flow.SYNTH_GAINS)
sd_h = weights.synth_state_dict(hift.param_shapes(), 1986, hift.SYNTH_GAINS)
m = B200CosyVoice2Model(precision="fp32", device=0, workspace_gb=4.0)
m.load_state_dicts(sd_l, sd_f, sd_h)
text, ptext, ptok, U = cases.lm_case()
_, _, pfeat, emb = cases.flow_case(P=9)
pfeat = pfeat[:, :18]
g = np.load("tests/golden/stream_tts.npz")
print("U", tuple(U.shape), "text", tuple(text.shape), "ptext", tuple(ptext.shape), "ptok", tuple(ptok.shape))
ids_o = lm.inference(sd_l, text, ptext, ptok, U, num_layers=NL)
print("oracle ids", len(ids_o))
for sps in (32, 8):
    ids = m.lm_generate([text], [ptext], [ptok], uniforms=U[:, None, :], steps_per_sync=sps)[0]
    nd = next((i for i, (a, b) in enumerate(zip(ids, ids_o)) if a != b), None)
    print(f"gpu ids sps={sps}: n={len(ids)} first diff at {nd}")
tok = torch.tensor(ids_o, dtype=torch.int32).unsqueeze(0)
mel_o = flow.inference(sd_f, tok, ptok, pfeat, emb, cfg=cfg)            # [1,80,T]
mel, lens = m.flow_batch([tok], [ptok], [pfeat], [emb]); m.stream.synchronize()
mel_o_tm = mel_o[0].t().contiguous()
print("mel", tuple(mel.shape), lens, "oracle", tuple(mel_o_tm.shape), "max|d|", (mel.cpu() - mel_o_tm).abs().max().item(),
      "per-100-frame", [(mel.cpu()[i:i + 100] - mel_o_tm[i:i + 100]).abs().max().item() for i in range(0, mel.shape[0], 100)])
noise = stream_noise(0, mel_o_tm.shape[0] * 480)
wav_o, src_o = hift.inference(sd_h, mel_o, noise)
wav, src = m.hift_batch(mel_o_tm.to(m.device), [mel_o_tm.shape[0]], noise=noise.to(m.device)); m.stream.synchronize()
wav_o = torch.as_tensor(wav_o).reshape(-1); wav = wav.cpu().reshape(-1)
print("wav (oracle mel) max|d|", (wav - wav_o).abs().max().item(), "per-second",
      [(wav[i:i + 24000] - wav_o[i:i + 24000]).abs().max().item() for i in range(0, wav.numel(), 24000)])
print("src max|d|", (src.cpu().reshape(-1) - torch.as_tensor(src_o).reshape(-1)).abs().max().item())
print("oracle wav vs golden", np.abs(wav_o.numpy() - g["offline_wav"].reshape(-1)).max())
