"""CPU, world_size 2 over gloo: the batch-shard + final-gather plumbing of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent_diffusion_speech_b200.distributed import gather_mels, shard_bounds, sharded_infer


def test_shard_bounds_cover_batch():
    for n in (1, 2, 7, 64, 512):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


class _FakeModel:
    """Stands in for Unit2Mel.forward: a per-utterance function of (units, spk_id, noise)."""

    def __call__(self, units, volume, spk_id=None, infer=True, noise=None, **kw):
        return units[..., :4] * 2.0 + spk_id.float()[:, :, None] + noise[:, 0, :4, :].transpose(1, 2)


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        units = torch.randn(n_items, 6, 16, generator=g)
        spk = torch.randint(1, 9, (n_items, 1), generator=g)
        noise = torch.randn(n_items, 1, 8, 6, generator=g)
        out = sharded_infer(_FakeModel(), units, spk, noise=noise)
        want = _FakeModel()(units, None, spk_id=spk, noise=noise)
        ret[rank] = bool(torch.equal(out, want))
        lo, hi = shard_bounds(n_items, world, rank)
        part = gather_mels(want[lo:hi], n_items)
        ret[rank] = ret[rank] and bool(torch.equal(part, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5])
def test_sharded_infer_equals_unsharded_world2(n_items):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, n_items, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
