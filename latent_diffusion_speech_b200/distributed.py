"""Batch sharding across the GPUs of one box.

Utterances are independent — GroupNorm/LayerNorm/attention are per sample and every solver
coefficient is batch invariant — so the batch is split contiguously across ranks, weights are
replicated, and there is NO collective inside the step loop.  The only exchange is one final
``all_gather`` of the ``[B/G, T, out_dims]`` mel shards (NCCL over NVLink 5 / NVSwitch on the GPU
box; the same code runs over ``gloo`` in the CPU tests).  The reference has no inference-time
parallelism at all (tools/infer_tools.py:14 is single-device).

``sharded_train_loss`` is the data-parallel form of the training-loss FORWARD (``forward(infer=False)``): every rank evaluates the
squared-error sum of its shard and ONE ``all_reduce`` of two scalars (sum, count) gives the global mean — the loss a data-parallel
trainer would log (the reference's DDP path, 20_train_diffusion.py, averages gradients; there is no backward pass in this package).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank's utterances; the first n_items % world_size ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_mels(mel_shard: torch.Tensor, n_items: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gathers ragged [b_r, T, M] shards into [n_items, T, M] on every rank (rank order = batch order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mel_shard
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_items, world, r) for r in range(world)]
    max_b = max(hi - lo for lo, hi in sizes)
    pad = mel_shard
    if mel_shard.shape[0] < max_b:
        pad = torch.zeros((max_b,) + tuple(mel_shard.shape[1:]), dtype=mel_shard.dtype, device=mel_shard.device)
        pad[: mel_shard.shape[0]] = mel_shard
    out = torch.empty((world * max_b,) + tuple(mel_shard.shape[1:]), dtype=mel_shard.dtype, device=mel_shard.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.view(world, max_b, *mel_shard.shape[1:])
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def sharded_infer(model, units: torch.Tensor, spk_id: Optional[torch.Tensor], *, noise: Optional[torch.Tensor] = None,
                  gt_spec: Optional[torch.Tensor] = None, step_noise=None, gather: bool = True,
                  group: Optional[dist.ProcessGroup] = None, out_dims: Optional[int] = None, **forward_kwargs) -> torch.Tensor:
    """Runs ``model(units[lo:hi], ...)`` for this rank's slice of the GLOBAL batch and gathers the mels.

    Every per-utterance input is given for the global batch and sliced here (nothing is scattered): ``units``,
    ``spk_id``, ``noise`` ([B,1,M,T]), ``gt_spec`` ([B,T,M], shallow diffusion) and ``step_noise`` (callable
    ``(j0, j1) -> [j1-j0, B, 1, M, T]``, DDPM).  Per-utterance noise makes the result independent of the number of ranks.
    A rank whose shard is empty (fewer utterances than ranks) skips the model call and contributes zero rows
    (``out_dims`` = mel bins of those rows; default ``model.decoder.out_dims``)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = units.shape[0]
    for name, t in (("spk_id", spk_id), ("noise", noise), ("gt_spec", gt_spec)):
        if t is not None and t.shape[0] != n:
            raise ValueError(f"{name} has {t.shape[0]} rows, expected the global batch size {n}")
    lo, hi = shard_bounds(n, world, rank)
    if hi > lo:
        local_step_noise = None
        if step_noise is not None:
            local_step_noise = lambda j0, j1: step_noise(j0, j1)[:, lo:hi]
        mel = model(units[lo:hi], None, spk_id=None if spk_id is None else spk_id[lo:hi], infer=True,
                    noise=None if noise is None else noise[lo:hi], gt_spec=None if gt_spec is None else gt_spec[lo:hi],
                    step_noise=local_step_noise, **forward_kwargs)
    else:
        if out_dims is None:
            out_dims = model.decoder.out_dims
        mel = torch.zeros((0, units.shape[1], out_dims), dtype=torch.float32, device=units.device)
    return gather_mels(mel, n, group) if gather else mel


def sharded_train_loss(model, units: torch.Tensor, spk_id: Optional[torch.Tensor], gt_spec: torch.Tensor, *,
                       t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                       group: Optional[dist.ProcessGroup] = None, **forward_kwargs) -> torch.Tensor:
    """Global-batch mean of the diffusion loss: ``model(units[lo:hi], ..., gt_spec=gt_spec[lo:hi], infer=False)`` on this rank's slice,
    weighted by its element count, then one ``all_reduce(SUM)`` of (weighted loss, count).  All inputs are given for the GLOBAL batch
    (``t`` [B] int64 and ``noise`` [B,1,M,T] make the result independent of the number of ranks); a rank with an empty shard
    contributes (0, 0).  Returns a 0-dim tensor equal on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = units.shape[0]
    for name, x in (("spk_id", spk_id), ("gt_spec", gt_spec), ("t", t), ("noise", noise)):
        if x is not None and x.shape[0] != n:
            raise ValueError(f"{name} has {x.shape[0]} rows, expected the global batch size {n}")
    lo, hi = shard_bounds(n, world, rank)
    acc = torch.zeros(2, dtype=torch.float64, device=units.device)
    if hi > lo:
        loss = model(units[lo:hi], None, spk_id=None if spk_id is None else spk_id[lo:hi], gt_spec=gt_spec[lo:hi], infer=False,
                     t=None if t is None else t[lo:hi], noise=None if noise is None else noise[lo:hi], **forward_kwargs)
        count = float((hi - lo) * gt_spec.shape[1] * gt_spec.shape[2])
        acc[0], acc[1] = loss.double() * count, count
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return (acc[0] / acc[1]).float()
