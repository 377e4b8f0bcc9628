"""Host-side multi-GPU logic on CPU: world_size-2 gloo (-m "not gpu")."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from watersurfacerendering_b200 import sharding as S


def test_round_robin_partition_is_exact():
    for n in (0, 1, 7, 64, 256, 1000):
        for world in (1, 2, 4, 8):
            parts = [S.shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([], np.int64)
            assert allidx.tolist() == list(range(n))
            assert [len(p) for p in parts] == S.shard_counts(n, world)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        S.shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = S.shard_indices(n_items, rank, world)
    # stand-in for the per-frame amplitude each rank's GPU would return
    local = [float(1000 + 3 * i) for i in idx]
    full = S.gather_in_global_order(local, n_items, rank, world)
    # timing reduction as bench.py does it: max over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, full.tolist(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 64])
def test_two_rank_gloo_gather(n_items):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [float(1000 + 3 * i) for i in range(n_items)]
    for rank, full, tmax in res:
        assert full == want
        assert tmax == 11.0
