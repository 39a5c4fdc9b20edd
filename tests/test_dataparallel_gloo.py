"""CPU, world_size 2 over gloo: the data-parallel host logic of the hot path -- batch sharding
by rank and the single gradient exchange (in-place mean all-reduce of a flat buffer plus coalesced loose
tensors).  The same functions run over NCCL in bench.py / engine.GraphedTrainStep."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from object_detection_destr_b200.dataparallel import allreduce_mean_, shard_range
        import bench
        # per-rank "gradients": a flat buffer with parameter views into it, plus two loose tensors and a None
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(1000, generator=g)
        views = [flat[:600].view(20, 30), flat[600:]]
        loose = [torch.randn(7, 3, generator=g), None, torch.randn(5, generator=g)]
        mine = [flat.clone(), loose[0].clone(), loose[2].clone()]
        allreduce_mean_([flat], loose, world=world)
        # expected: recompute every rank's tensors locally
        exp = [torch.zeros_like(t) for t in mine]
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            for e, shape in zip(exp, ((1000,), (7, 3), (5,))):
                e += torch.randn(*shape, generator=gr)
        exp = [e / world for e in exp]
        ok = torch.allclose(flat, exp[0], atol=1e-6) and torch.allclose(loose[0], exp[1], atol=1e-6) \
            and torch.allclose(loose[2], exp[2], atol=1e-6)
        ok &= torch.equal(views[0], flat[:600].view(20, 30))  # parameter views still alias the reduced buffer
        # all ranks hold identical results
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        ok &= all(torch.equal(gathered[0], t) for t in gathered)
        # sharding: disjoint, covering, and the synthetic batches of different ranks differ
        shards = [list(shard_range(16, r, world)) for r in range(world)]
        ok &= sorted(sum(shards, [])) == list(range(16))
        b0, b1 = bench.make_batch(0, 0, 2), bench.make_batch(1, 0, 2)
        ok &= not torch.equal(b0[0], b1[0])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_and_sharding_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_range_rejects_ragged():
    sys.path.insert(0, ROOT)
    from object_detection_destr_b200.dataparallel import shard_range
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    assert list(shard_range(8, 1, 2)) == [4, 5, 6, 7]
