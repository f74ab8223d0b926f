"""GPU experiment (not a test): ONE denoiser evaluation at the headline shape (B=64, T=864) through the C ABI —
a short target for ncu launch lists (405 launches instead of the 65k of a bench run).
usage: gpu_nfe_once.py [precision] [B] [T] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from latent_diffusion_speech_b200.unit2mel import Unit2Mel  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 864
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(1234)
model = Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval().cuda().set_precision(precision)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 128, T, device="cuda", generator=g)
cond = torch.randn(B, T, 256, device="cuda", generator=g)
for _ in range(reps):
    eps = model.denoise(x, cond, 500.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eps = model.denoise(x, cond, 500.0)
e1.record()
torch.cuda.synchronize()
print(f"one NFE B={B} T={T} {precision}: {e0.elapsed_time(e1):.3f} ms, launches so far {model._engine.kernel_launches}, finite {bool(torch.isfinite(eps).all())}")
