"""TEST INFRASTRUCTURE ONLY — CPU restatement of the HiFi-VAEGAN ``Generator`` decode (latent / mel frames -> waveform), the
step right after the Unit2Mel sampling path.  Never imported by the product package.

Follows, in the reference tree:
  encoder/hifi_vaegan/modules/models.py:224-266   Generator.__init__ / forward
  encoder/hifi_vaegan/modules/models.py:161-200   ResBlock1 (three (dilated conv, conv) pairs with residuals)
  encoder/hifi_vaegan/modules/models.py:203-221   ResBlock2 (two dilated convs with residuals)
  encoder/hifi_vaegan/modules/commons.py:13-14    get_padding
  encoder/hifi_vaegan/hifi_vaegan.py:52-65        Hifi_VAEGAN.forward (transpose [B,T,C] -> [B,C,T], Generator, remove_weight_norm)
  diffusion/vocoder.py:31-32                      Vocoder.infer

Pinned by tests/test_vocoder_oracle.py: torch.equal against the UNMODIFIED reference ``Generator`` imported from
/root/reference (authoring container) and against tests/golden/vocoder_*.npz produced by it (oracle/make_golden_vocoder.py).

The generator configuration ``h`` lives inside the vocoder checkpoint (``decoder.pth["config"]``, hifi_vaegan.py:6-8), which the
reference tree does not ship; DEFAULT_H is the HiFi-GAN V1 layout whose upsample rates multiply to the hop size (512) the
reference hard-codes (hifi_vaegan.py:20) with ``inter_channels`` = Unit2Mel's out_dims (128).
"""
from typing import Dict

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1          # models.py:12
DEFAULT_H = {
    "sampling_rate": 44100, "hop_size": 512, "inter_channels": 128, "upsample_initial_channel": 512,
    "upsample_rates": [8, 8, 4, 2], "upsample_kernel_sizes": [16, 16, 8, 4], "resblock": "1",
    "resblock_kernel_sizes": [3, 7, 11], "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
}


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``remove_weight_norm`` (models.py:258-264) as a pure function of the state dict: w = g * v / ||v|| over all dims but 0
    (torch.nn.utils.weight_norm, dim=0), both for the legacy ``weight_g/weight_v`` and the parametrization key names."""
    out = {}
    for k, v in sd.items():
        for g_key, v_key in (("weight_g", "weight_v"), ("parametrizations.weight.original0", "parametrizations.weight.original1")):
            if k.endswith(v_key):
                g = sd[k[: -len(v_key)] + g_key]
                out[k[: -len(v_key)] + "weight"] = torch._weight_norm(v, g, 0)      # the op torch's own remove_weight_norm evaluates
        if not (k.endswith("weight_g") or k.endswith("weight_v") or ".parametrizations." in k):
            out[k] = v
    return out


def generator_forward(sd: Dict[str, torch.Tensor], h: dict, z: torch.Tensor) -> torch.Tensor:
    """z [B, inter_channels, T] -> wav [B, 1, T * prod(upsample_rates)]   (models.py:248-256; weights without weight norm)."""
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(z, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u + 1) // 2)
        xs = None
        for j, (ks, dil) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            r = _resblock(sd, f"resblocks.{i * nk + j}", h["resblock"], x, ks, dil)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)                          # default slope 0.01 here (models.py:252), not LRELU_SLOPE
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def _resblock(sd, key, kind, x, ks, dil):
    if kind == "1":
        for n, d in enumerate(dil):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = F.conv1d(xt, sd[f"{key}.convs1.{n}.weight"], sd[f"{key}.convs1.{n}.bias"], dilation=d, padding=get_padding(ks, d))
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = F.conv1d(xt, sd[f"{key}.convs2.{n}.weight"], sd[f"{key}.convs2.{n}.bias"], padding=get_padding(ks, 1))
            x = xt + x
        return x
    for n, d in enumerate(dil[:2]):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{key}.convs.{n}.weight"], sd[f"{key}.convs.{n}.bias"], dilation=d, padding=get_padding(ks, d))
        x = xt + x
    return x


def vocoder_infer(sd: Dict[str, torch.Tensor], h: dict, mel: torch.Tensor) -> torch.Tensor:
    """Vocoder.infer(mel [B, T, C]) -> wav [B, 1, T * hop]   (diffusion/vocoder.py:31-32, hifi_vaegan.py:52-65)."""
    return generator_forward(fold_weight_norm(sd), h, mel.transpose(-1, -2))


def synthetic_latents(B: int, T: int, C: int = 128, seed: int = 21) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, C, generator=g)
