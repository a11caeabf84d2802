"""GPU-box timing of the device preprocess (Resize + CenterCrop, then the fused ToTensor + Normalize ingest) on a
ragged batch of ImageNet-like image sizes: images/s of resize_crop_u8 alone and of the whole score_images chain,
next to Pillow / torchvision on one host core."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import synth  # noqa: E402
from mcm_b200.engine import McmEngine  # noqa: E402

cfg = synth.CFGS["ViT-B/16"]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=256)
eng.set_text_bank(synth.synth_unit_bank(1000, cfg.proj, 3), already_unit=True)
rng = np.random.default_rng(0)
sizes = [(375, 500), (500, 375), (333, 500), (500, 333), (480, 640), (600, 800), (768, 1024), (256, 256)]
imgs = [rng.integers(0, 256, size=(*sizes[i % len(sizes)], 3), dtype=np.uint8) for i in range(256)]
mb = sum(a.size for a in imgs) / 1e6
for _ in range(2):
    eng.resize_crop_u8(imgs)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    out = eng.resize_crop_u8(imgs)
torch.cuda.synchronize()
t_resize = (time.perf_counter() - t0) / reps
# source already packed on the device: host planning + launch (CPU time of the call) and the kernel (CUDA events)
sizes_b = np.array([a.size for a in imgs], dtype=np.int64)
offs = np.zeros(len(imgs), dtype=np.int64)
offs[1:] = np.cumsum(sizes_b[:-1])
src = torch.from_numpy(np.concatenate([a.reshape(-1) for a in imgs])).cuda()
hs = np.array([a.shape[0] for a in imgs], dtype=np.int32)
ws = np.array([a.shape[1] for a in imgs], dtype=np.int32)
dst = torch.empty((len(imgs), 224, 224, 3), dtype=torch.uint8, device="cuda")
eng.resize_crop_u8_packed(src, offs, hs, ws, out=dst)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    eng.resize_crop_u8_packed(src, offs, hs, ws, out=dst)
e1.record()
t_call = (time.perf_counter() - t0) / reps
torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) * 1e-3 / reps
eng.score_images(imgs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    s = eng.score_images(imgs)
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / reps
# one-call host stream from a pinned packed buffer: pipelined H2D + device preprocess + scoring
packed_host = torch.from_numpy(np.concatenate([a.reshape(-1) for a in imgs] * 4)).pin_memory()
offs4 = np.concatenate([offs + k * int(sizes_b.sum()) for k in range(4)])
hs4, ws4 = np.tile(hs, 4), np.tile(ws, 4)
eng.score_stream_host_images(packed_host, offs4[:512], hs4[:512], ws4[:512], batch=256)
t0 = time.perf_counter()
eng.score_stream_host_images(packed_host, offs4, hs4, ws4, batch=256)
t_stream = time.perf_counter() - t0
rec = dict(n=len(imgs), source_mb=mb, host_stream_images_per_s=len(offs4) / t_stream, host_stream_h2d_gbs=4 * mb / 1e3 / t_stream, packed_call_cpu_ms=1e3 * t_call, packed_device_images_per_s=len(imgs) / t_dev,
           packed_device_source_gbs=mb / 1e3 / t_dev, resize_crop_images_per_s=len(imgs) / t_resize, score_images_per_s=len(imgs) / t_all)
try:
    from PIL import Image
    import torchvision.transforms as T
    tf = T.Compose([T.Resize(224), T.CenterCrop(224)])
    pil = [Image.fromarray(a) for a in imgs[:64]]
    t0 = time.perf_counter()
    for im in pil:
        np.asarray(tf(im))
    rec["pillow_one_core_images_per_s"] = len(pil) / (time.perf_counter() - t0)
except Exception as e:  # noqa: BLE001
    rec["pillow"] = repr(e)
print(json.dumps(rec))
