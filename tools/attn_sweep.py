"""GPU-box timing of the attention kernel alone (ViT-B/16 shape): 20 launches, CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import _lib, synth  # noqa: E402
from mcm_b200.engine import McmEngine  # noqa: E402

if os.environ.get("SWEEP_LIB"):       # A/B variant built by mcm_b200.build.build_variant
    _lib.use_library(os.environ["SWEEP_LIB"])

cfg = synth.CFGS["tiny"]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=4)
shapes = [tuple(int(v) for v in c.split(",")) for c in os.environ.get("SWEEP_SHAPES", "256,197,12;256,50,12;64,197,12;128,257,16").split(";")]
for b, S, H in shapes:
    qkv = (torch.randn(b * S, 3 * H * 64, device="cuda") * 1.5).to(torch.float16)
    for _ in range(3):
        out = eng.dbg_attention(qkv, b, S, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.dbg_attention(qkv, b, S, H)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    # spot check of the first two images against torch fp32 on the same fp16 q, k, v
    nb = min(b, 2)
    qq, kk, vv = [t.float().reshape(nb, S, H, 64).transpose(1, 2) for t in qkv[: nb * S].split(H * 64, dim=1)]
    ref = (torch.softmax(qq @ kk.transpose(-1, -2) * 0.125, dim=-1) @ vv).transpose(1, 2).reshape(nb * S, H * 64)
    err = float((out[: nb * S].float() - ref).abs().max())
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("SWEEP_LIB", "default")), b=b, S=S, H=H, us=us,
                          tflops=4.0 * b * H * S * S * 64 / us / 1e6, max_abs_err_vs_torch=err)), flush=True)
