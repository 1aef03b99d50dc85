"""N>1 host logic on CPU: world_size-2 gloo process group, image sharding and the max-over-ranks
timing reduce bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nquant_android_b200.sharding import shard_range, max_over_ranks


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(1024, rank, world)
    owned = torch.zeros(1024, dtype=torch.int64)
    owned[b:e] = 1
    dist.all_reduce(owned)                      # test-only collective: every image owned exactly once
    t = max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    q.put((rank, int(owned.min()), int(owned.max()), t, e - b))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:4] for r in res] == [(1, 1, 11.0), (1, 1, 11.0)]
    assert sum(r[4] for r in res) == 1024
