"""N > 1 host logic on CPU: contiguous stream sharding + the score all-gather, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcm_b200 import parallel


def test_shard_bounds_cover_stream():
    for n in (0, 1, 7, 8, 9, 50000, 5640):
        for W in (1, 2, 3, 4, 8):
            got = []
            for r in range(W):
                lo, hi = parallel.shard_bounds(n, r, W)
                assert 0 <= lo <= hi <= n and hi - lo <= parallel.shard_len(n, W)
                got.extend(range(lo, hi))
            assert got == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = np.arange(n, dtype=np.float32) * 0.5 - 3
        lo, hi = parallel.shard_bounds(n, rank, world)
        out = parallel.gather_scores(full[lo:hi], n)
        ok = out.dtype == np.float32 and out.shape == (n,) and np.array_equal(out, full)
        bad = False
        if n > 1:   # a wrong-sized slice is rejected on every rank BEFORE the collective is entered
            try:
                parallel.gather_scores(full[lo:hi + 1] if hi < n else full[lo:hi][:-1], n)
            except ValueError:
                bad = True
        q.put((rank, bool(ok), bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 8, 1])
def test_gather_scores_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert all(r[2] for r in res if n > 1)


def test_single_process_passthrough():
    out = parallel.gather_scores(torch.arange(5, dtype=torch.float32), 5)
    assert out.tolist() == [0, 1, 2, 3, 4]
