"""mcm_allgather_scores: the score collation of a sharded stream through the C ABI on a raw
ncclComm_t -- one rank on any GPU box, two ranks when the box has two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_allgather_single_rank(engine_factory):
    from mcm_b200 import parallel
    eng, _, _ = engine_factory("tiny", 5, 8)
    comm = parallel.NcclComm(device=0)
    try:
        local = torch.arange(37, dtype=torch.float32, device="cuda") * 0.25 - 2
        out = parallel.gather_scores_nccl(eng, comm, local, 37)
        assert out.dtype == np.float32 and np.array_equal(out, local.cpu().numpy())
        with pytest.raises(ValueError):
            eng.allgather_scores(0, local, local.clone())          # NULL communicator
    finally:
        comm.close()


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from mcm_b200 import parallel, synth
    from mcm_b200.engine import McmEngine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.CFGS["tiny"]
        eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=16, device=rank)
        eng.set_text_bank(synth.synth_unit_bank(10, cfg.proj, 3))
        comm = parallel.NcclComm(device=rank)
        imgs = torch.from_numpy(synth.synth_images(n, 3))
        lo, hi = parallel.shard_bounds(n, rank, world)
        with torch.cuda.device(rank):
            local = torch.cat([eng.score(imgs[s:min(s + 16, hi)].cuda(rank)).clone() for s in range(lo, hi, 16)]) if hi > lo \
                else torch.empty(0, device=f"cuda:{rank}")
            got = parallel.gather_scores_nccl(eng, comm, local, n)
            full = torch.cat([eng.score(imgs[s:s + 16].cuda(rank)).clone() for s in range(0, n, 16)]).cpu().numpy()
        q.put((rank, bool(np.array_equal(got, full))))
        comm.close()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [37, 2])
def test_allgather_two_ranks(n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
