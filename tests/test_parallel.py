"""CPU tier: the N>1 path (shard assemblies, broadcast the index image once, gather hits) on gloo, world_size 2."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from kaptive_b200 import parallel


def test_shard_range_partitions():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = np.arange(100_003, dtype=np.uint64).view(np.uint8) if rank == 0 else None
        got = parallel.broadcast_bytes(img, 0, device="cpu")
        ok_b = bool(np.array_equal(got, np.arange(100_003, dtype=np.uint64).view(np.uint8)))
        lo, hi = parallel.shard_range(11, rank, world)
        hits = {"asm_id": np.arange(hi - lo, dtype=np.int32), "gene": np.full(hi - lo, rank, dtype=np.int32)}
        allh = parallel.gather_hits(hits, lo, 0)
        if rank == 0:
            q.put((ok_b, allh["asm_id"].tolist(), allh["gene"].tolist()))
        else:
            q.put((ok_b, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_and_gather_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=100) for _ in ps]
    for p in ps:
        p.join(30)
    assert all(r[0] for r in res)
    full = [r for r in res if r[1] is not None][0]
    assert full[1] == list(range(11))
    assert full[2] == [0] * 6 + [1] * 5
