"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  python oracle/make_golden.py

Each fixture stores the seeds/shapes that regenerate the inputs (``synthetic_inputs``), the
reference's output mel, and a checksum of the random-init weights (``torch.manual_seed(1234)``
default init of the reference ``Unit2Mel(1280, 323, 128, 2, [256,384,512,512], 8, 256, 1.0)``),
so a consumer can prove it rebuilt the very same parameters without shipping 475 MB of them.
The reference draws its noise with ``torch.randn``; we inject the per-utterance seeded noise of
``synthetic_inputs`` by temporarily replacing ``torch.randn``/``randn_like`` (call order:
initial noise / q_sample noise first, then one draw per DDPM iteration, diffusion.py:207,170,118).
"""
import contextlib
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unit2mel_oracle as O  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 1234

# name, B, T, method, infer_speedup, k_step (None = full 1000), n DDPM noises
CASES = [
    ("dpm20_b2_t40", 2, 40, "dpm-solver", 50, None, 0),
    ("dpm20_b1_t37", 1, 37, "dpm-solver", 50, None, 0),       # 8 does not divide T -> forced upsample size
    ("dpm8_b1_t24", 1, 24, "dpm-solver", 125, None, 0),        # steps < 10 -> lower_order_final
    ("unipc10_b2_t37", 2, 37, "unipc", 100, None, 0),
    ("unipc10_b2_t48", 2, 48, "unipc", 100, None, 0),
    ("shallow_dpm20_b2_t32", 2, 32, "dpm-solver", 5, 100, 0),
    ("shallow_unipc10_b2_t32", 2, 32, "unipc", 10, 100, 0),
    ("ddpm12_b2_t24", 2, 24, None, 1, 12, 12),                 # shallow start + 12 ancestral steps
    ("ddim20_b2_t40", 2, 40, "ddim", 50, None, 0),
    ("shallow_ddim10_b2_t37", 2, 37, "ddim", 10, 100, 0),
    ("pndm20_b1_t40", 1, 40, "pndm", 50, None, 0),             # the reference's PLMS path only runs for B == 1
    ("shallow_pndm10_b1_t37", 1, 37, "pndm", 10, 100, 0),
    ("dpm20_b1_t432", 1, 432, "dpm-solver", 50, None, 0),       # BASELINE configs[0]: one ~5 s utterance, 20-step DPM-Solver
]


def state_dict_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@contextlib.contextmanager
def inject_randn(noises):
    queue = [n for n in noises]
    orig, orig_like = torch.randn, torch.randn_like

    def fake(*a, **k):
        return queue.pop(0).clone()

    torch.randn, torch.randn_like = fake, (lambda x, **k: fake())
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig, orig_like


def main():
    ref = import_reference()
    torch.manual_seed(WEIGHT_SEED)
    model = ref.Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval()
    sd = model.state_dict()
    csum = state_dict_checksum(sd)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    with torch.no_grad():
        for name, B, T, method, speedup, k_step, n_noise in CASES:
            if os.path.exists(os.path.join(GOLDEN_DIR, name + ".npz")) and "--force" not in sys.argv:
                continue
            units, spk, noise, steps, gt = O.synthetic_inputs(B, T, n_step_noises=n_noise, gt=k_step is not None)
            with inject_randn([noise] + steps):
                if k_step is None:
                    mel = model(units, None, spk_id=spk, infer=True, infer_speedup=speedup, method=method)
                else:   # shallow diffusion is only reachable through the decoder
                    cond = model.unit_embed(units) + model.spk_embed(spk - 1)
                    mel = model.decoder(cond, gt_spec=gt, infer=True, infer_speedup=speedup, method=method,
                                        k_step=k_step)
            np.savez_compressed(
                os.path.join(GOLDEN_DIR, name + ".npz"),
                B=B, T=T, method=method or "", infer_speedup=speedup, k_step=-1 if k_step is None else k_step,
                n_step_noises=n_noise, weight_seed=WEIGHT_SEED, weights_sha256=csum,
                mel=mel.numpy(), torch_version=torch.__version__)
            print(name, tuple(mel.shape), float(mel.abs().max()))

        # single denoiser evaluations (lds_denoise parity): eps for fractional and integer timesteps
        for name, B, T, t in [("nfe_b2_t40_t999", 2, 40, 999.0), ("nfe_b1_t37_t417p25", 1, 37, 417.25),
                              ("nfe_b2_t64_t0", 2, 64, 0.0)]:
            if os.path.exists(os.path.join(GOLDEN_DIR, name + ".npz")) and "--force" not in sys.argv:
                continue
            units, spk, noise, _, _ = O.synthetic_inputs(B, T)
            cond = (model.unit_embed(units) + model.spk_embed(spk - 1)).transpose(1, 2)
            inp = torch.cat([noise[:, 0], cond], dim=-2)
            eps = model.decoder.denoise_fn(inp, torch.full((B,), t)).sample
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), B=B, T=T, t=t, weight_seed=WEIGHT_SEED,
                                weights_sha256=csum, eps=eps.numpy(), cond=cond.numpy(),
                                torch_version=torch.__version__)
            print(name, tuple(eps.shape), float(eps.abs().max()))

        # training-loss forward (Unit2Mel.forward(infer=False), unit2mel.py:73-89 -> diffusion.py:193-201 -> p_losses :173-187): the
        # reference's own forward with its randint (timesteps) and randn_like (noise) draws replaced by seeded ones
        for name, B, T, tvals in [("trainloss_b2_t40", 2, 40, [17, 903]), ("trainloss_b3_t37", 3, 37, [0, 999, 412])]:
            if os.path.exists(os.path.join(GOLDEN_DIR, name + ".npz")) and "--force" not in sys.argv:
                continue
            units, spk, noise, _, gt = O.synthetic_inputs(B, T, gt=True)
            tt = torch.tensor(tvals, dtype=torch.long)
            orig_randint = torch.randint
            torch.randint = lambda *a, **k: tt.clone()
            try:
                with inject_randn([noise]):
                    loss = model(units, None, spk_id=spk, gt_spec=gt, infer=False)
            finally:
                torch.randint = orig_randint
            cond = (model.unit_embed(units) + model.spk_embed(spk - 1)).transpose(1, 2)
            spec = model.decoder.norm_spec(gt).transpose(1, 2)[:, None]
            loss_l1 = model.decoder.p_losses(spec, tt, cond=cond, noise=noise, loss_type="l1")
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), B=B, T=T, t=np.array(tvals), weight_seed=WEIGHT_SEED,
                                weights_sha256=csum, loss_l2=float(loss), loss_l1=float(loss_l1), torch_version=torch.__version__)
            print(name, "l2", float(loss), "l1", float(loss_l1))


if __name__ == "__main__":
    main()
