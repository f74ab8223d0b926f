"""CPU, world_size 2 over gloo: the batch-shard + final-gather plumbing of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent_diffusion_speech_b200.distributed import gather_mels, shard_bounds, sharded_infer, sharded_train_loss


def test_shard_bounds_cover_batch():
    for n in (1, 2, 7, 64, 512):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


class _FakeModel:
    """Stands in for Unit2Mel.forward: a per-utterance function of (units, spk_id, noise, gt_spec, step_noise)."""

    def __call__(self, units, volume, spk_id=None, infer=True, noise=None, gt_spec=None, step_noise=None, **kw):
        out = units[..., :4] * 2.0 + spk_id.float()[:, :, None] + noise[:, 0, :4, :].transpose(1, 2)
        if gt_spec is not None:
            assert gt_spec.shape[0] == units.shape[0]
            out = out + gt_spec[..., :4]
        if step_noise is not None:
            z = step_noise(0, 3)                        # [3, B_local, 1, M, T]
            assert z.shape[1] == units.shape[0]
            out = out + z.sum(0)[:, 0, :4, :].transpose(1, 2)
        return out


def _worker(rank, world, port, n_items, shallow, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        units = torch.randn(n_items, 6, 16, generator=g)
        spk = torch.randint(1, 9, (n_items, 1), generator=g)
        noise = torch.randn(n_items, 1, 8, 6, generator=g)
        gt = torch.randn(n_items, 6, 8, generator=g) if shallow else None
        steps = torch.randn(3, n_items, 1, 8, 6, generator=g)
        sn = (lambda j0, j1: steps[j0:j1]) if shallow else None
        out = sharded_infer(_FakeModel(), units, spk, noise=noise, gt_spec=gt, step_noise=sn, out_dims=4)
        want = _FakeModel()(units, None, spk_id=spk, noise=noise, gt_spec=gt, step_noise=sn)
        ret[rank] = bool(torch.equal(out, want))
        lo, hi = shard_bounds(n_items, world, rank)
        part = gather_mels(want[lo:hi], n_items)
        ret[rank] = ret[rank] and bool(torch.equal(part, want))
    finally:
        dist.destroy_process_group()


# n_items = 1 < world: rank 1 has an empty shard (must not call the model, must still take part in the gather)
@pytest.mark.parametrize("n_items,shallow", [(4, False), (5, False), (5, True), (1, True)])
def test_sharded_infer_equals_unsharded_world2(n_items, shallow):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, n_items, shallow, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_sharded_infer_rejects_mismatched_global_inputs():
    units = torch.randn(4, 6, 16)
    with pytest.raises(ValueError):
        sharded_infer(_FakeModel(), units, torch.ones(4, 1, dtype=torch.long), noise=torch.randn(4, 1, 8, 6), gt_spec=torch.randn(3, 6, 8))


class _FakeLossModel:
    """Stands in for Unit2Mel.forward(infer=False): the mean of a per-element function over the local shard."""

    def __call__(self, units, volume, spk_id=None, gt_spec=None, infer=False, t=None, noise=None, **kw):
        assert not infer and gt_spec.shape[0] == units.shape[0] == t.shape[0] == noise.shape[0]
        err = noise[:, 0].transpose(1, 2) - gt_spec * t.float()[:, None, None] * 1e-3 + spk_id.float()[:, :, None]
        return (err * err).mean()


def _loss_worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        units = torch.randn(n_items, 6, 16, generator=g)
        spk = torch.randint(1, 9, (n_items, 1), generator=g)
        gt = torch.randn(n_items, 6, 8, generator=g)
        t = torch.randint(0, 1000, (n_items,), generator=g)
        noise = torch.randn(n_items, 1, 8, 6, generator=g)
        got = sharded_train_loss(_FakeLossModel(), units, spk, gt, t=t, noise=noise)
        want = _FakeLossModel()(units, None, spk_id=spk, gt_spec=gt, t=t, noise=noise)
        ret[rank] = bool(abs(float(got) - float(want)) <= 1e-6 * abs(float(want)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5, 1])
def test_sharded_train_loss_equals_global_mean_world2(n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_loss_worker, args=(2, port, n_items, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
