"""Multi-process (world_size 2, gloo, CPU) tests of the data-parallel plumbing: batch sharding and the flat-gradient
all-reduce that the modules' backward issues (slotdiffusion_b200/parallel.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from slotdiffusion_b200 import parallel
    try:
        assert not parallel.enabled()
        parallel.enable_grad_allreduce()
        assert parallel.enabled()
        # the flat buffer of a rank holds "its" gradients; after the reduce every rank holds the mean
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        parallel.allreduce_flat(flat)
        parallel.wait_all()
        expect = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok = torch.allclose(flat, expect)
        # sharding covers the global batch exactly once, remainders to the first ranks
        spans = [parallel.shard_batch(65, r, world) for r in range(world)]
        ok = ok and spans[0][0] == 0 and spans[-1][1] == 65 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        ok = ok and parallel.shard_batch(64) == (rank * 32, rank * 32 + 32)
        # data-parallel equivalence on a toy loss: mean over ranks of per-shard mean gradients == full-batch gradient
        torch.manual_seed(0)
        w = torch.randn(8, requires_grad=True)
        x = torch.randn(64, 8)
        a, b = parallel.shard_batch(64)
        (x[a:b] @ w).square().mean().backward()
        g = w.grad.clone()
        parallel.allreduce_flat(g)
        w2 = w.detach().clone().requires_grad_(True)
        (x @ w2).square().mean().backward()
        ok = ok and torch.allclose(g, w2.grad, atol=1e-6)
        parallel.disable_grad_allreduce()
        ok = ok and not parallel.enabled()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
