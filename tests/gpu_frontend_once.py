"""GPU experiment (not a test): ONE pass of the two stages either side of the sampler — a short target for ncu launch lists.
  units  : log-mel + whisper-width AudioEncoder (1280 x 20 heads, N layers) + alignment, B x 30 s of audio
  vocoder: HiFi-VAEGAN generator decode, B x T frames
usage: gpu_frontend_once.py units|vocoder [B] [layers|T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

what = sys.argv[1] if len(sys.argv) > 1 else "units"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(1234)
if what == "units":
    from latent_diffusion_speech_b200 import units as UN
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    enc = UN.AudioEncoder(128, 1280, 20, layers).eval().cuda()
    audio = (0.1 * torch.randn(B, 480000, device="cuda")).clamp(-1, 1)
    for _ in range(2):
        mel = UN.log_mel_spectrogram(audio)
        u = UN.units_forced_alignment(enc(mel), n_frames=2583)
    torch.cuda.synchronize()
    print("units", tuple(u.shape), "launches", enc._engine.kernel_launches, "finite", bool(torch.isfinite(u).all()))
else:
    from latent_diffusion_speech_b200.vocoder import Vocoder
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 864
    voc = Vocoder("hifi-vaegan", None, device="cuda")
    mel = torch.randn(B, T, voc.dimension, device="cuda")
    for _ in range(2):
        wav = voc.infer(mel)
    torch.cuda.synchronize()
    print("vocoder", tuple(wav.shape), "launches", voc.generator._engine.kernel_launches, "finite", bool(torch.isfinite(wav).all()))
