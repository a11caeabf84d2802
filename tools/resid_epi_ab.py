"""Timing A/B of the residual epilogues at the ViT-B/16 shapes (batch 512): fp32 residual stream (round 1) vs the fp16
(hi, lo) pair, TMA form (K = 768, out_proj) and LSU form (K = 3072, fc2).  CUDA events, 10 launches each."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import synth  # noqa: E402
if os.environ.get("AB_LIB"):          # an A/B build made by mcm_b200.build.build_variant
    from mcm_b200 import _lib
    _lib.use_library(os.environ["AB_LIB"])
from mcm_b200.engine import McmEngine  # noqa: E402

cfg = synth.CFGS["tiny"]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=4)
M, N = 512 * 197, 768
only = os.environ.get("AB_ONLY", "")
for K in (768, 3072):
    a = torch.randn(M, K, device="cuda").to(torch.float16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.float16)
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    x_hi = x.to(torch.float16)
    x_lo = (x - x_hi.float()).to(torch.float16)
    runs = {"fp32": lambda: eng.dbg_gemm_resid_ln(a, w, bias, x, in_place=True, mutate=True),
            "h2": lambda: eng.dbg_gemm_resid_h2(a, w, bias, x_hi, x_lo)}
    for name, fn in runs.items():
        if only and only != f"{name}{K}":
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(json.dumps(dict(lib=os.path.basename(os.environ.get("AB_LIB", "default")), kind=name, M=M, N=N, K=K, us=us, tflops=2.0 * M * N * K / us / 1e6)), flush=True)
