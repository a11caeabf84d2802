"""Minimal driver for ncu: build the ViT-B/16 engine, warm up, then run N steps of the hot path
between cudaProfilerStart/Stop (use ncu --profile-from-start off).  Not a benchmark."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", default="ViT-B/16")
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--no-shortcut", action="store_true")
ap.add_argument("--lib", default="", help="an A/B build made by mcm_b200.build.build_variant")
a = ap.parse_args()

from mcm_b200 import _lib, synth
if a.lib:
    _lib.use_library(os.path.abspath(a.lib))
from mcm_b200.engine import McmEngine

cfg = synth.CFGS[a.model]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=a.batch)
eng.set_text_bank(synth.synth_unit_bank(a.K, cfg.proj, 3), already_unit=True)
if a.no_shortcut:
    eng.set_cls_shortcut(False)
x = torch.randn(a.batch, 3, 224, 224, device="cuda")
for _ in range(2):
    eng.score(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    eng.score(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
