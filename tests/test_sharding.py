"""Host logic of the multi-GPU path: contiguous env shards, no collective on the
solve path, optional gather -- exercised with world_size-2 gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import biped_pympc_b200 as bp


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for ws in (1, 2, 3, 8):
            spans = [bp.shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        bp.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 12, dtype=torch.float32).view(n, 2, 6)
        solver = bp.ShardedSolver(n, solve_fn=lambda ins, rot: ins[0] * 2.0)
        local = solver.local([full])
        w = solver.solve(local, None)
        assert w.shape[0] == solver.hi - solver.lo
        g = solver.gather(w)
        ok = torch.equal(g, full * 2.0)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_two_rank_shard_and_gather_gloo(n):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
