"""GPU-box timing sweep of the tcgen05 GEMM: separates the k-loop rate from the per-tile epilogue cost.

Writes gpurun_out/gemm_sweep.jsonl: one line per (M, N, K, epi) with the average of 20 launches
(CUDA events), TFLOP/s, cycles per tile and per k-block at the nominal 1.965 GHz."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import synth  # noqa: E402
from mcm_b200.engine import McmEngine  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "gemm_sweep.jsonl")
M = int(os.environ.get("SWEEP_M", "50432"))
CASES = [(768, 64, 0), (768, 64, 1), (768, 64, 2), (3072, 64, 0), (3072, 64, 1),
         (768, 768, 0), (768, 768, 2), (768, 3072, 0), (768, 3072, 2), (768, 6144, 0), (256, 6144, 0),
         (2304, 768, 0), (3072, 768, 0), (3072, 768, 1), (3072, 1536, 0)]

if os.environ.get("SWEEP_CASES"):   # e.g. "3072,768,0;768,768,6"
    CASES = [tuple(int(v) for v in c.split(",")) for c in os.environ["SWEEP_CASES"].split(";")]
TAG = os.environ.get("SWEEP_TAG", "")
if os.environ.get("SWEEP_LIB"):     # an A/B build made by mcm_b200.build.build_variant
    from mcm_b200 import _lib
    _lib.use_library(os.path.abspath(os.environ["SWEEP_LIB"]))
cfg = synth.CFGS["tiny"]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=4)
os.makedirs(os.path.dirname(OUT), exist_ok=True)
g = torch.Generator(device="cuda").manual_seed(0)
for N, K, epi in CASES:
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi in (2, 6) else None
    if epi in (4, 5):     # LayerNorm-folded projection: K is the row length the statistics cover
        stats = torch.randn(2 * (K // 256) if K % 256 == 0 else K // 64, M, 2, device="cuda", generator=g).abs() + 1.0
        stats[..., 1] += stats[..., 0] ** 2
        cvec = torch.randn(N, device="cuda", generator=g)
        run = lambda: eng.dbg_gemm_ln(a, w, bias, cvec, stats, K, gelu=(epi == 5))
    elif epi == 6:
        run = lambda: eng.dbg_gemm_resid_ln(a, w, bias, resid, in_place=True, mutate=True)
    else:
        run = lambda: eng.dbg_gemm(a, w, bias, resid, epi)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    tiles = ((M + 255) // 256) * (N // 256)
    per_cluster = -(-tiles // 74)
    cyc_tile = us * 1e-6 * 1.965e9 / per_cluster
    rec = dict(tag=TAG, N=N, K=K, epi=epi, us=us, tflops=2.0 * M * N * K / us / 1e6, tiles=tiles, tiles_per_cluster=per_cluster,
               cycles_per_tile=cyc_tile, cycles_per_kblock=cyc_tile / (K // 64))
    print(json.dumps(rec), flush=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")
    del a, w, resid
