"""Drop-in ``Unit2Mel`` whose inference path runs on the B200-native CUDA library.

Mirrors the reference module (diffusion/unit2mel.py:10-89): same constructor, same
``forward`` signature (plus the trailing optional ``k_step`` the reference's declared
interface promises but never wired), same ``state_dict()`` keys and shapes, same
``load_model_vocoder`` / ``load_svc_model`` helpers and the same ``config.yaml`` schema.

What changed underneath: ``forward(infer=True)`` uploads the parameters once to an
``lds_handle`` (C ABI, include/lds_b200.h) and runs conditioning, every denoiser evaluation and
every solver update as hand-written sm_100a kernels.  There is no PyTorch fallback for that path.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import yaml

from .capi import Engine
from .denoiser_params import DenoiserParams
from .diffusion import GaussianDiffusion

ENCODER_OUT_CHANNELS = {"whisper_large_v3": 1280, "contentvec768l12": 768, "xlsr_53_56k": 1024}


def get_encoder_out_channels(encoder: str) -> int:
    """tools/tools.py:257-263 (the only symbol of that module the hot path needs)."""
    if encoder not in ENCODER_OUT_CHANNELS:
        raise ValueError(f"unknown units encoder {encoder!r}")
    return ENCODER_OUT_CHANNELS[encoder]


class DotDict(dict):
    def __getattr__(*args):
        val = dict.get(*args)
        return DotDict(val) if type(val) is dict else val

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


class Unit2Mel(nn.Module):
    def __init__(self, input_channel, n_spk, out_dims=128, n_layers=2, block_out_channels=(256, 384, 512, 512),
                 n_heads=8, n_hidden=256, acoustic_scale=1.0):
        super().__init__()
        self.unit_embed = nn.Linear(input_channel, n_hidden)
        self.aug_shift_embed = None
        self.volume_embed = None
        self.is_tts = True            # the reference reads this attribute without ever setting it (unit2mel.py:74)
        self.n_spk = n_spk
        if n_spk is not None and n_spk > 1:
            self.spk_embed = nn.Embedding(n_spk, n_hidden)
        self.decoder = GaussianDiffusion(
            DenoiserParams(in_channels=out_dims + n_hidden, out_channels=out_dims,
                           block_out_channels=block_out_channels, layers_per_block=n_layers, n_heads=n_heads,
                           norm_groups=8),
            out_dims=out_dims, acoustic_scale=acoustic_scale)
        self._hp = dict(input_channel=input_channel, n_spk=n_spk, out_dims=out_dims, n_layers=n_layers,
                        block_out_channels=tuple(block_out_channels), n_heads=n_heads, n_hidden=n_hidden,
                        acoustic_scale=acoustic_scale)
        self.precision = "fp32"       # "fp32": split-f16 tcgen05 (fp32-accurate); "bf16": bf16 tcgen05; "fp32_ffma": CUDA-core fp32
        self._engine: Optional[Engine] = None
        self.decoder._engine_provider = self._get_engine
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_engine())

    # ---- engine management ------------------------------------------------------------------
    def invalidate_engine(self) -> None:
        """Call after mutating parameters in place; load_state_dict()/to() do it automatically."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        if getattr(self, "decoder", None) is not None:
            self.decoder.invalidate_programs()  # sampler programs are functions of the (checkpointed) schedule buffers

    def _apply(self, fn, *a, **k):
        self.invalidate_engine()
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str) -> "Unit2Mel":
        if precision not in ("fp32", "bf16", "fp32_ffma"):
            raise ValueError(precision)
        if precision != self.precision:
            self.precision = precision
            self.invalidate_engine()
        return self

    def _get_engine(self, device: torch.device) -> Engine:
        device = torch.device(device)
        if self._engine is None or self._engine.device != device or self._engine.precision != self.precision:
            self.invalidate_engine()
            eng = Engine(device=device, precision=self.precision, **self._hp)
            eng.load_state_dict(self.state_dict())
            self._engine = eng
        return self._engine

    # ---- reference interface ----------------------------------------------------------------
    def forward(self, units, volume, spk_id=None, aug_shift=None, gt_spec=None, infer=True, infer_speedup=10,
                method="unipc", use_tqdm=False, k_step=None, noise=None, step_noise=None, t=None):
        """units [B,T,input_channel], spk_id [B,1] int64 -> mel [B,T,out_dims] (unit2mel.py:73-89).  ``infer=False`` returns the training
        loss of ``p_losses`` (diffusion.py:173-201) as a 0-dim tensor computed FORWARD-only on the library — the validation loss of
        diffusion/solver.py:56-62; there is no autograd graph, backward and optimiser steps stay with the reference (``t`` / ``noise``
        optionally replace its ``randint`` / ``randn_like`` draws)."""
        if not infer:
            if gt_spec is None:
                raise ValueError("infer=False needs gt_spec")
            if not (volume is None or self.is_tts):
                raise NotImplementedError("volume_embed is None in the reference (unit2mel.py:56); pass volume=None")
            if not units.is_cuda:
                raise RuntimeError("Unit2Mel runs on a CUDA device only (no CPU fallback): move inputs to cuda")
            eng = self._get_engine(units.device)
            with torch.no_grad():
                b, t_frames, _ = units.shape
                self.decoder.prepare(eng, b, t_frames, "dpm-solver", max(2, self.decoder.k_step // 2), self.decoder.k_step)
                cond = eng.cond(units, spk_id if (self.n_spk is not None and self.n_spk > 1) else None)
                return self.decoder(cond, gt_spec=gt_spec, infer=False, k_step=k_step, noise=noise, t=t)
        if not (volume is None or self.is_tts):
            raise NotImplementedError("volume_embed is None in the reference (unit2mel.py:56); pass volume=None")
        if not units.is_cuda:
            raise RuntimeError("Unit2Mel inference runs on a CUDA device only (no CPU fallback): move inputs to cuda")
        eng = self._get_engine(units.device)
        b, t_frames, _ = units.shape
        # cond = unit_embed(units) + 0 + spk_embed(spk_id - 1); aug_shift_embed is None (unit2mel.py:55,84)
        shallow = gt_spec is not None and k_step is not None
        self.decoder.prepare(eng, b, t_frames, method, infer_speedup, int(k_step) if shallow else self.decoder.k_step)
        cond = eng.cond(units, spk_id if (self.n_spk is not None and self.n_spk > 1) else None)
        return self.decoder(cond, gt_spec=gt_spec, infer=True, infer_speedup=infer_speedup, method=method,
                            k_step=k_step, use_tqdm=use_tqdm, noise=noise, step_noise=step_noise)

    @torch.no_grad()
    def denoise(self, x, cond, t):
        """One denoiser evaluation eps(x[B,M,T] | cond[B,T,H], t) — parity hook for UNet1DConditionModel.forward."""
        from .sampler_tables import timestep_sinusoid
        eng = self._get_engine(x.device)
        b, _, t_frames = x.shape
        if (eng.B, eng.T) != (b, t_frames):
            eng.plan(b, t_frames, 0, None, None, key=None)
        row = timestep_sinusoid(torch.tensor([float(t)]), self._hp["block_out_channels"][0]).numpy()[0]
        return eng.denoise(x, cond, row)


def load_svc_model(args, vocoder_dimension):
    """unit2mel.py:37-49 with the declared 8-parameter constructor (the reference passes a stray
    ``use_pitch_aug`` positional and raises TypeError)."""
    return Unit2Mel(
        get_encoder_out_channels(args["data"]["encoder"]),
        args["common"]["n_spk"],
        vocoder_dimension,
        args["diffusion"]["model"]["n_layers"],
        args["diffusion"]["model"]["block_out_channels"],
        args["diffusion"]["model"]["n_heads"],
        args["diffusion"]["model"]["n_hidden"],
        args["data"]["acoustic_scale"],
    )


def load_model_vocoder(model_path, device="cpu", loaded_vocoder=None, vocoder_factory=None):
    """unit2mel.py:18-35: reads config.yaml next to the checkpoint, builds the vocoder (through
    ``vocoder_factory(type, ckpt, device)`` — the vocoder stays on the reference PyTorch path and is
    not part of this package), builds the model and loads ``ckpt['model']`` strictly."""
    config_file = os.path.join(os.path.split(model_path)[0], "config.yaml")
    with open(config_file, "r") as config:
        args = DotDict(yaml.safe_load(config))
    if loaded_vocoder is not None:
        vocoder = loaded_vocoder
    elif vocoder_factory is not None:
        vocoder = vocoder_factory(args["common"]["vocoder"]["type"], args["common"]["vocoder"]["ckpt"], device)
    else:
        from diffusion.vocoder import Vocoder   # reference class, resolved from the caller's PYTHONPATH
        vocoder = Vocoder(args["common"]["vocoder"]["type"], args["common"]["vocoder"]["ckpt"], device=device)
    model = load_svc_model(args=args, vocoder_dimension=vocoder.dimension)
    ckpt = torch.load(model_path, map_location=torch.device(device))
    model.to(device)
    model.load_state_dict(ckpt["model"])
    model.eval()
    return model, vocoder, args
