"""Parameter container for the 1-D conditional U-Net denoiser.

This module holds *parameters only*.  It reproduces the reference denoiser's
``state_dict()`` key names, shapes and default-initialisation order so that

* ``load_state_dict(ckpt['model'])`` of a reference checkpoint works strictly, and
* ``torch.manual_seed(s); Unit2Mel(...)`` yields bit-identical random-init weights
  to the reference class built under the same seed (tests pin this with a checksum).

No arithmetic lives here: the forward pass of the denoiser is executed by the CUDA
library (``csrc/``) through the C ABI in ``include/lds_b200.h``.

Reference structure followed (file:line, relative to the reference tree):
  diffusion/unet1d/unet_1d_condition.py:151-607   (top-level wiring)
  diffusion/unet1d/unet_1d_blocks.py:516-623      (mid block)
  diffusion/unet1d/unet_1d_blocks.py:861-1096     (down blocks)
  diffusion/unet1d/unet_1d_blocks.py:1985-2206    (up blocks)
  diffusion/unet1d/resnet.py:104-223,461-590      (up/down samplers, resnet)
  diffusion/unet1d/transformer_1d.py:65-190       (transformer wrapper)
  diffusion/unet1d/attention.py:46-128,206-301    (transformer block, GEGLU FF)
  diffusion/unet1d/attention_processor.py:40-143  (attention projections)
  diffusion/unet1d/embeddings.py:157-186          (timestep MLP)
"""
from __future__ import annotations

from typing import Sequence

import torch.nn as nn

TIME_EMBED_MULT = 4          # time_embed_dim = 4 * block_out_channels[0]
RESNET_EPS = 1e-5            # unet_1d_condition.py:174 (norm_eps)
TRANSFORMER_GN_EPS = 1e-6    # transformer_1d.py:134
LAYERNORM_EPS = 1e-5         # nn.LayerNorm default, attention.py:83


class ResnetParams(nn.Module):
    """norm1, conv1, time_emb_proj, norm2, conv2[, conv_shortcut] (resnet.py:527-590)."""

    def __init__(self, c_in: int, c_out: int, temb: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, c_in, eps=RESNET_EPS)
        self.conv1 = nn.Conv1d(c_in, c_out, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, 2 * c_out)       # scale_shift conditioning
        self.norm2 = nn.GroupNorm(groups, c_out, eps=RESNET_EPS)
        self.conv2 = nn.Conv1d(c_out, c_out, 3, padding=1)
        if c_in != c_out:
            self.conv_shortcut = nn.Conv1d(c_in, c_out, 1)


class AttentionParams(nn.Module):
    """to_q/to_k/to_v without bias, to_out.0 with bias (attention_processor.py:128-143)."""

    def __init__(self, dim: int):
        super().__init__()
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])


class _GegluProj(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, 2 * inner)


class FeedForwardParams(nn.Module):
    """net.0.proj (C -> 8C), net.2 (4C -> C) (attention.py:224-249)."""

    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GegluProj(dim, 4 * dim), nn.Identity(), nn.Linear(4 * dim, dim)])


class TransformerBlockParams(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LAYERNORM_EPS)
        self.attn1 = AttentionParams(dim)
        self.norm2 = nn.LayerNorm(dim, eps=LAYERNORM_EPS)
        self.attn2 = AttentionParams(dim)
        self.norm3 = nn.LayerNorm(dim, eps=LAYERNORM_EPS)
        self.ff = FeedForwardParams(dim)


class TransformerParams(nn.Module):
    """norm (GroupNorm eps 1e-6), proj_in (1x1), transformer_blocks.0, proj_out (1x1)."""

    def __init__(self, dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=TRANSFORMER_GN_EPS)
        self.proj_in = nn.Conv1d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([TransformerBlockParams(dim)])
        self.proj_out = nn.Conv1d(dim, dim, 1)


class _Resample(nn.Module):
    def __init__(self, dim: int, stride: int):
        super().__init__()
        self.conv = nn.Conv1d(dim, dim, 3, stride=stride, padding=1)


class DownBlockParams(nn.Module):
    def __init__(self, c_in, c_out, temb, groups, n_layers, with_attention, with_downsample):
        super().__init__()
        resnets, attentions = [], []
        for i in range(n_layers):                      # creation order = RNG order
            resnets.append(ResnetParams(c_in if i == 0 else c_out, c_out, temb, groups))
            if with_attention:
                attentions.append(TransformerParams(c_out, groups))
        if with_attention:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        if with_downsample:
            self.downsamplers = nn.ModuleList([_Resample(c_out, 2)])


class MidBlockParams(nn.Module):
    def __init__(self, dim, temb, groups):
        super().__init__()
        r0 = ResnetParams(dim, dim, temb, groups)
        a0 = TransformerParams(dim, groups)
        r1 = ResnetParams(dim, dim, temb, groups)
        self.attentions = nn.ModuleList([a0])
        self.resnets = nn.ModuleList([r0, r1])


class UpBlockParams(nn.Module):
    def __init__(self, c_skip_last, c_prev, c_out, temb, groups, n_layers, with_attention, with_upsample):
        super().__init__()
        resnets, attentions = [], []
        for i in range(n_layers):
            skip = c_skip_last if i == n_layers - 1 else c_out
            c_in = c_prev if i == 0 else c_out
            resnets.append(ResnetParams(c_in + skip, c_out, temb, groups))
            if with_attention:
                attentions.append(TransformerParams(c_out, groups))
        if with_attention:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        if with_upsample:
            self.upsamplers = nn.ModuleList([_Resample(c_out, 1)])


class _TimeMLP(nn.Module):
    def __init__(self, c_in, c_t):
        super().__init__()
        self.linear_1 = nn.Linear(c_in, c_t)
        self.linear_2 = nn.Linear(c_t, c_t)


class DenoiserParams(nn.Module):
    """Parameters of the U-Net: 3 attention down blocks + 1 plain, mid, 1 plain up + 3 attention."""

    def __init__(self, in_channels: int, out_channels: int, block_out_channels: Sequence[int],
                 layers_per_block: int = 2, n_heads: int = 8, norm_groups: int = 8):
        super().__init__()
        ch = list(block_out_channels)
        n = len(ch)
        temb = TIME_EMBED_MULT * ch[0]
        self.in_channels, self.out_channels = in_channels, out_channels
        self.block_out_channels, self.layers_per_block = ch, layers_per_block
        self.n_heads, self.norm_groups = n_heads, norm_groups

        self.conv_in = nn.Conv1d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = _TimeMLP(ch[0], temb)

        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()      # registered before mid_block, as in the reference
        c_out = ch[0]
        for i in range(n):
            c_in, c_out = c_out, ch[i]
            last = i == n - 1
            self.down_blocks.append(DownBlockParams(c_in, c_out, temb, norm_groups, layers_per_block,
                                                    with_attention=not last, with_downsample=not last))
        self.mid_block = MidBlockParams(ch[-1], temb, norm_groups)

        rev = ch[::-1]
        c_out = rev[0]
        for i in range(n):
            c_prev, c_out = c_out, rev[i]
            c_skip_last = rev[min(i + 1, n - 1)]
            last = i == n - 1
            self.up_blocks.append(UpBlockParams(c_skip_last, c_prev, c_out, temb, norm_groups,
                                                layers_per_block + 1, with_attention=i > 0,
                                                with_upsample=not last))
        self.conv_norm_out = nn.GroupNorm(norm_groups, ch[0], eps=RESNET_EPS)
        self.conv_out = nn.Conv1d(ch[0], out_channels, 3, padding=1)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("DenoiserParams holds parameters only; the forward pass runs in the CUDA library")
