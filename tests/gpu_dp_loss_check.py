"""GPU experiment under torchrun (not a pytest test): the data-parallel training-loss forward — `sharded_train_loss` over NCCL on N
ranks against the single-GPU loss of the global batch on rank 0 and against the fp64 oracle (B = 8 x T = 216, seeded t / noise).
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/gpu_dp_loss_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from latent_diffusion_speech_b200.distributed import sharded_train_loss  # noqa: E402
from latent_diffusion_speech_b200.unit2mel import Unit2Mel  # noqa: E402
from oracle import unit2mel_oracle as O  # noqa: E402   (checker only)

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
torch.manual_seed(1234)
model = Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval()
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.cuda()
B, T = 8, 216
units, spk, noise, _, gt = O.synthetic_inputs(B, T, gt=True, seed=9)
t = torch.tensor([0, 999, 17, 503, 250, 750, 64, 901])
dp = sharded_train_loss(model, units.cuda(), spk.cuda(), gt.cuda(), t=t, noise=noise.cuda())
ok = True
if rank == 0:
    full = model(units.cuda(), None, spk_id=spk.cuda(), gt_spec=gt.cuda(), infer=False, t=t, noise=noise.cuda())
    with torch.no_grad():
        l64 = O.unit2mel_train_loss(sd, O.DEFAULT_CFG, units, spk, gt, t, noise, "l2", dtype=torch.float64)
    r1, r2 = abs(float(dp) - float(full)) / float(full), abs(float(dp) - float(l64)) / float(l64)
    ok = r1 <= 1e-6 and r2 <= 2e-6
    print(f"dp loss over {world} ranks {float(dp):.8f} | single GPU {float(full):.8f} (rel {r1:.2e}) | fp64 oracle {float(l64):.8f} (rel {r2:.2e}) | {'OK' if ok else 'FAIL'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
