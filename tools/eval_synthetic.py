#!/usr/bin/env python
"""End-to-end MCM evaluation on synthetic streams -- the shape of ``eval_ood_detection.py:main``
(``:53-99``: ID scores, a loop over OOD sets, AUROC / AUPR / FPR95 per set and their mean) driven
through the drop-in API, on 1..8 GPUs (BASELINE configs 3-5).

    python tools/eval_synthetic.py --model ViT-B/16 --K 1000 --n-id 5000 --ood 10000
    torchrun --nproc-per-node 8 tools/eval_synthetic.py --model ViT-L/14 --K 1000 \
        --n-id 50000 --ood 10000,10000,10000,5640

Every rank scores the contiguous slice ``parallel.shard_bounds`` of each stream with replicated
weights and bank; one all-gather per stream collates the scores; rank 0 prints one JSON line.
Streams are generated on the device in slabs (prototype harness of mcm_b200.synth: ID = prototype +
noise, OOD = fresh noise), so nothing but scores crosses PCIe.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ViT-B/16")
    ap.add_argument("--K", type=int, default=1000)
    ap.add_argument("--n-id", type=int, default=5000)
    ap.add_argument("--ood", default="10000", help="comma-separated OOD stream sizes")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--noise", type=float, default=0.8)
    ap.add_argument("--T", type=float, default=1.0)
    ap.add_argument("--score", default="MCM")
    a = ap.parse_args()

    import torch.distributed as dist
    from mcm_b200 import metrics, parallel, synth
    from mcm_b200.engine import McmEngine

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CFGS[a.model]
    eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=a.batch, device=local)

    # prototype bank: K seeded prototype images -> their features -> centred unit rows
    g = torch.Generator(device=dev).manual_seed(100)
    protos = torch.randn((a.K, 3, cfg.image_size, cfg.image_size), device=dev, generator=g)
    feats = torch.cat([eng.image_features(protos[s:s + a.batch]) for s in range(0, a.K, a.batch)]).cpu().numpy()
    eng.set_text_bank(synth.centred_prototype_bank(feats))

    def score_stream(n, seed, is_id):
        """Images are defined per GLOBAL slab of `batch` images (seeded by the slab index), so the stream --
        and therefore every score and metric -- is identical for any number of ranks."""
        lo, hi = parallel.shard_bounds(n, rank, world)
        out = torch.empty((hi - lo,), dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev)
        for k in range(lo // a.batch, -(-hi // a.batch) if hi > lo else 0):
            s0 = k * a.batch
            gen.manual_seed(seed * 1_000_003 + k)
            x = torch.randn((a.batch, 3, cfg.image_size, cfg.image_size), device=dev, generator=gen)
            if is_id:
                idx = (torch.arange(s0, s0 + a.batch, device=dev) % a.K)
                x = x * a.noise + protos[idx]
            else:
                x = x * float(np.sqrt(1.0 + a.noise ** 2))
            u, v = max(lo, s0), min(hi, s0 + a.batch)          # part of the slab owned by this rank
            eng.score(x[u - s0:v - s0].contiguous(), T=a.T, score=a.score, out=out[u - lo:v - lo])
        return parallel.gather_scores(out, n)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    import hashlib
    digest = hashlib.sha1()
    in_score = score_stream(a.n_id, 1, True)
    digest.update(in_score.tobytes())
    rows = []
    for j, n in enumerate(int(v) for v in a.ood.split(",")):
        out_score = score_stream(n, 10 + j, False)
        digest.update(out_score.tobytes())
        auroc, aupr, fpr = metrics.get_measures(-in_score, -out_score)      # utils/detection_util.py:259
        rows.append(dict(n_ood=n, auroc=auroc, aupr=aupr, fpr95=float(fpr)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        n_total = a.n_id + sum(r["n_ood"] for r in rows)
        print(json.dumps(dict(model=a.model, K=a.K, n_gpus=world, n_id=a.n_id, sets=rows,
                              mean_auroc=float(np.mean([r["auroc"] for r in rows])),
                              mean_fpr95=float(np.mean([r["fpr95"] for r in rows])),
                              score_sha1=digest.hexdigest(),   # identical for any number of ranks
                              images=n_total, seconds=dt, images_per_s=n_total / dt,
                              note="wall clock incl. on-device stream generation; bench.py is the timing instrument")))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
