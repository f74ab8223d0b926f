"""HiFi-VAEGAN decode on the B200 library: ``Vocoder.infer(mel)`` (diffusion/vocoder.py:5-32) ->
``Hifi_VAEGAN.forward`` (encoder/hifi_vaegan/hifi_vaegan.py:52-65) -> ``Generator.forward``
(encoder/hifi_vaegan/modules/models.py:224-266), i.e. the step right after ``Unit2Mel.forward``.

``Generator`` keeps the reference's constructor argument (the ``h`` dictionary stored in the vocoder checkpoint), its
``state_dict()`` keys / shapes in the inference form (after ``remove_weight_norm``) and its default random init, so that a
reference checkpoint — with or without weight-norm parametrisation — loads, and ``torch.manual_seed(s); Generator(h)`` holds
bit-identical parameters to the reference's.  ``forward`` runs on the CUDA library (csrc/vocoder.cu); there is no CPU fallback.
The VAE *encoder* (``Vocoder.extract``, audio -> latent) is not on this path and stays with the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from .capi import LdsError, load_library

VOC_MAX = 8
DEFAULT_H = {
    # the generator configuration lives in decoder.pth["config"] (hifi_vaegan.py:6-8), which the reference tree does not ship:
    # HiFi-GAN V1 layout whose upsample rates multiply to the hop size the reference hard-codes (512, hifi_vaegan.py:20)
    "sampling_rate": 44100, "hop_size": 512, "inter_channels": 128, "upsample_initial_channel": 512,
    "upsample_rates": [8, 8, 4, 2], "upsample_kernel_sizes": [16, 16, 8, 4], "resblock": "1",
    "resblock_kernel_sizes": [3, 7, 11], "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
}


class VocoderConfig(C.Structure):
    _fields_ = [("inter_channels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("n_ups", C.c_int32),
                ("upsample_rates", C.c_int32 * VOC_MAX), ("upsample_kernel_sizes", C.c_int32 * VOC_MAX),
                ("resblock_kind", C.c_int32), ("n_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * VOC_MAX),
                ("resblock_dilations", (C.c_int32 * 3) * VOC_MAX)]


def _pad(k: int, d: int = 1) -> int:
    return int((k * d - d) / 2)


def _draw_like_init_weights(m: nn.Module) -> None:
    """The reference applies ``init_weights`` (commons.py:5-11) to weight-normed convolutions: ``m.weight.data.normal_(0, 0.01)``
    writes the *derived* weight, which the next forward recomputes from (g, v) — the call changes nothing but the RNG stream.
    Consuming the same draws keeps every later default init identical to the reference's."""
    for mod in m.modules():
        if "Conv" in mod.__class__.__name__:
            mod.weight.data.normal_(0.0, 0.01)


class _ResBlock(nn.Module):
    def __init__(self, kind: str, channels: int, k: int, dilation):
        super().__init__()
        mk = lambda d: weight_norm(nn.Conv1d(channels, channels, k, 1, dilation=d, padding=_pad(k, d)))
        if kind == "1":
            self.convs1 = nn.ModuleList([mk(d) for d in dilation[:3]])
            _draw_like_init_weights(self.convs1)
            self.convs2 = nn.ModuleList([mk(1) for _ in dilation[:3]])
            _draw_like_init_weights(self.convs2)
        else:
            self.convs = nn.ModuleList([mk(d) for d in dilation[:2]])
            _draw_like_init_weights(self.convs)


class Generator(nn.Module):
    """Parameter container + CUDA forward with the reference Generator's interface (models.py:224-266)."""

    def __init__(self, h: dict):
        super().__init__()
        self.h = dict(h)
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        c0 = h["upsample_initial_channel"]
        self.conv_pre = weight_norm(nn.Conv1d(h["inter_channels"], c0, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            self.ups.append(weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u + 1) // 2)))
        self.resblocks = nn.ModuleList()
        for i in range(len(self.ups)):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
                self.resblocks.append(_ResBlock(str(h["resblock"]), ch, k, d))
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        _draw_like_init_weights(self.ups)
        _draw_like_init_weights(self.conv_post)
        self.upp = 1
        for u in h["upsample_rates"]:
            self.upp *= int(u)
        self._weight_norm_removed = False
        self.remove_weight_norm()                 # inference form: plain `weight` parameters, the keys the C library loads
        self._engine = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # ---- reference interface ------------------------------------------------------------------------------------------
    def remove_weight_norm(self):
        if self._weight_norm_removed:
            return
        for mod in self.modules():
            if isinstance(mod, (nn.Conv1d, nn.ConvTranspose1d)) and hasattr(mod, "weight_g"):
                remove_weight_norm(mod)
        self._weight_norm_removed = True

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts the reference's checkpoints in either form: folds (weight_g, weight_v) pairs with the op torch's own
        ``remove_weight_norm`` evaluates, so that the parameters equal the reference's after hifi_vaegan.py:56-61."""
        sd = {}
        for k, v in state_dict.items():
            if k.endswith("weight_v"):
                sd[k[:-len("weight_v")] + "weight"] = torch._weight_norm(v, state_dict[k[:-len("weight_v")] + "weight_g"], 0)
            elif not k.endswith("weight_g"):
                sd[k] = v
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def _invalidate(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, inter_channels, T] -> wav [B, 1, T * hop] (models.py:248-256)."""
        return self.decode_frames(x.transpose(-1, -2))

    @torch.no_grad()
    def decode_frames(self, mel_btc: torch.Tensor) -> torch.Tensor:
        """mel [B, T, inter_channels] (the layout Unit2Mel returns) -> wav [B, 1, T * hop]."""
        if not mel_btc.is_cuda:
            raise RuntimeError("the vocoder runs on a CUDA device only (no CPU fallback): move the input to cuda")
        if self._engine is None or self._engine.device != mel_btc.device:
            self._invalidate()
            self._engine = VocoderEngine(self.h, mel_btc.device)
            self._engine.load_state_dict(self.state_dict())
        return self._engine.vocode(mel_btc)[:, None, :]


class VocoderEngine:
    """One ``lds_vocoder`` handle (C ABI, include/lds_b200.h)."""

    def __init__(self, h: dict, device: torch.device):
        if device.type != "cuda":
            raise LdsError("the vocoder runs on a CUDA device (B200, sm_100a) only; got %s" % device)
        self.lib = load_library()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        cfg = VocoderConfig()
        cfg.inter_channels, cfg.upsample_initial_channel = int(h["inter_channels"]), int(h["upsample_initial_channel"])
        cfg.n_ups, cfg.n_kernels = len(h["upsample_rates"]), len(h["resblock_kernel_sizes"])
        if cfg.n_ups > VOC_MAX or cfg.n_kernels > VOC_MAX:
            raise ValueError("too many upsample stages / resblock kernels")
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
        cfg.resblock_kind = 1 if str(h["resblock"]) == "1" else 2
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            cfg.resblock_kernel_sizes[j] = int(k)
            for n in range(3):
                cfg.resblock_dilations[j][n] = int(d[n]) if n < len(d) else 1
        self.cfg = cfg
        self.handle = C.c_void_p()
        self._check(self.lib.lds_vocoder_create(C.byref(cfg), self.index, C.byref(self.handle)), "lds_vocoder_create")
        self.hop = int(self.lib.lds_vocoder_hop(self.handle))

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise LdsError(f"{what} failed (status {rc}): {self.lib.lds_vocoder_last_error().decode(errors='replace')}")

    def load_state_dict(self, sd) -> None:
        for key, t in sd.items():
            t = t.detach().float().contiguous()
            shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
            self._check(self.lib.lds_vocoder_load_weight(self.handle, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), 0),
                        f"lds_vocoder_load_weight({key})")
        self._check(self.lib.lds_vocoder_finalize(self.handle), "lds_vocoder_finalize")

    def vocode(self, mel_btc: torch.Tensor) -> torch.Tensor:
        mel = mel_btc.to(device=self.device, dtype=torch.float32).contiguous()
        B, T, Cc = mel.shape
        if Cc != self.cfg.inter_channels:
            raise ValueError(f"mel has {Cc} bins, the generator expects {self.cfg.inter_channels}")
        wav = torch.empty(B, T * self.hop, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.index):
            self._check(self.lib.lds_vocode(self.handle, C.c_void_p(mel.data_ptr()), B, T, C.c_void_p(wav.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "lds_vocode")
        self._keep = mel
        return wav

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.lds_vocoder_launches(self.handle))

    @property
    def last_flops(self) -> float:
        return float(self.lib.lds_vocoder_last_flops(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.lds_vocoder_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Vocoder:
    """Mirror of the reference's ``Vocoder`` wrapper (diffusion/vocoder.py:5-32) for the decode direction."""

    def __init__(self, vocoder_type: str = "hifi-vaegan", vocoder_ckpt: Optional[str] = None, device=None, h: Optional[dict] = None):
        if vocoder_type != "hifi-vaegan":
            raise ValueError(f" [x] Unknown vocoder: {vocoder_type}")
        self.device = torch.device(device if device is not None else "cuda")
        self.vocoder_type = vocoder_type
        state = None
        if vocoder_ckpt is not None:                       # hifi_vaegan.py:5-8,56
            import os
            blob = torch.load(os.path.join(vocoder_ckpt, "decoder.pth"), map_location="cpu")
            h, state = blob["config"], blob["model"]
        self.h = dict(h or DEFAULT_H)
        self.generator = Generator(self.h).eval()
        if state is not None:
            self.generator.load_state_dict(state)
        self.generator.to(self.device)
        self.vocoder_sample_rate = self.h["sampling_rate"]
        self.vocoder_hop_size = self.h["hop_size"]
        self.dimension = self.h["inter_channels"]

    def extract(self, audio, sample_rate, keyshift=0, **kwargs):
        raise NotImplementedError("audio -> latent (the VAE encoder) is outside the B200 path; use the reference Vocoder.extract")

    def infer(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, n_frames, bins] -> wav [B, 1, n_frames * hop] (diffusion/vocoder.py:31-32)."""
        return self.generator.decode_frames(mel)
