"""World-size-2 gloo test (CPU) of the batch-sharding helpers: slices tile the batch, the gathered tensor is the
full batch on every rank, and the backward of the gather returns each rank the gradient of its own rows."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rayen_b200 import sharding


def test_shard_bounds_tile_the_batch():
    for batch in (0, 1, 7, 8, 262144):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batch, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        full = torch.randn(batch, 3, 1)                      # replicated "layer output" for the check
        w = torch.randn(batch, 3, 1)
        local = sharding.shard_batch(full).clone().requires_grad_(True)
        gathered = sharding.all_gather_outputs(local, batch=batch)
        ok_fwd = torch.equal(gathered, full)
        loss = (gathered * w).sum() / world                  # replicated loss, averaged over ranks
        loss.backward()
        lo, hi = sharding.shard_bounds(batch, rank, world)
        ok_bwd = torch.allclose(local.grad, w[lo:hi], atol=1e-6)
        results[rank] = (ok_fwd, ok_bwd, tuple(local.shape))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_all_gather_outputs_gloo_world2():
    for batch in (8, 7):
        manager = mp.Manager()
        results = manager.dict()
        mp.spawn(_worker, args=(2, _free_port(), batch, results), nprocs=2, join=True)
        assert len(results) == 2
        for rank in range(2):
            ok_fwd, ok_bwd, shape = results[rank]
            assert ok_fwd and ok_bwd, (batch, rank)
        assert results[0][2][0] + results[1][2][0] == batch
