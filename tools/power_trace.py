"""Power and SM clock of the step and of its kernels, each run back to back for a few seconds (GPU box).

The step runs at the board's power cap, so what a kernel costs is its ENERGY: this tool loops one workload at a
time (the whole forward, each GEMM shape of a ViT-B/16 layer, the attention kernel, cuBLAS for comparison),
samples NVML (instantaneous board power, SM clock, throttle reasons) every 20 ms from a side thread, and prints one
JSON line per workload: time per call, mean power and median SM clock over the settled part of the loop, energy per
call, TFLOP/s.  Not a benchmark (bench.py is); a reading aid for DESIGN.md section 4.

  python tools/power_trace.py [--seconds 3] [--batch 512] > gpurun_out/power_trace.jsonl
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcm_b200 import synth  # noqa: E402
from mcm_b200.engine import McmEngine  # noqa: E402

import pynvml  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--only", default="", help="comma-separated workload names")
ap.add_argument("--steps-only", default="", help="comma-separated model:precision pairs (e.g. ViT-L/14:fp16,ViT-B/16:split): "
                "loop only the whole step of each and exit")
a = ap.parse_args()

pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)


def power_w():
    try:
        v = pynvml.nvmlDeviceGetFieldValues(nv, [pynvml.NVML_FI_DEV_POWER_INSTANT])[0]
        if v.nvmlReturn == 0:
            return v.value.uiVal / 1000.0
    except Exception:
        pass
    return pynvml.nvmlDeviceGetPowerUsage(nv) / 1000.0


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows = []
        self.stop = False

    def run(self):
        while not self.stop:
            t = time.perf_counter()
            try:
                clk = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(nv)
            except Exception:
                clk, rs = 0, 0
            self.rows.append((t, power_w(), clk, rs))
            time.sleep(0.02)


def run_phase(name, fn, flops, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    chunk = max(1, int(0.05 / max(e0.elapsed_time(e1) * 1e-3, 1e-6)))   # ~50 ms of work between host syncs
    s = Sampler()
    s.start()
    t_start = time.perf_counter()
    settled_at, n_settled, ev0 = None, 0, None
    while True:
        now = time.perf_counter()
        if now - t_start >= seconds:
            break
        if settled_at is None and now - t_start >= min(1.0, seconds / 3):
            settled_at = now
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        for _ in range(chunk):
            fn()
        if settled_at is not None:
            n_settled += chunk
        torch.cuda.synchronize()
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    s.stop = True
    s.join()
    rows = [r for r in s.rows if settled_at is not None and settled_at <= r[0] <= t_end]
    ms = ev0.elapsed_time(ev1) / max(n_settled, 1)
    p = statistics.fmean(r[1] for r in rows) if rows else float("nan")
    reasons = 0
    for r in rows:
        reasons |= r[3]
    rec = dict(workload=name, ms_per_call=ms, calls=n_settled, power_w_mean=p, power_w_max=max((r[1] for r in rows), default=float("nan")),
               sm_mhz_median=statistics.median(r[2] for r in rows) if rows else 0, samples=len(rows),
               joule_per_call=p * ms * 1e-3, tflops=(flops / (ms * 1e-3) / 1e12) if flops else None,
               pj_per_flop=(p * ms * 1e-3 / flops * 1e12) if flops else None,
               sw_power_cap=bool(reasons & pynvml.nvmlClocksEventReasonSwPowerCap),
               thermal=bool(reasons & (pynvml.nvmlClocksEventReasonSwThermalSlowdown | pynvml.nvmlClocksEventReasonHwThermalSlowdown)))
    print(json.dumps(rec), flush=True)
    time.sleep(0.5)


if a.steps_only:
    for spec in a.steps_only.split(","):
        model, prec = spec.split(":")
        cfg = synth.CFGS[model]
        bb = a.batch if model != "ViT-L/14" else min(a.batch, 256)
        eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=bb)
        eng.set_text_bank(synth.synth_unit_bank(1000, cfg.proj, 3), already_unit=True)
        eng.set_precision(prec)
        xs = torch.randn(bb, 3, 224, 224, device="cuda")
        run_phase(f"step {model} batch {bb} precision {prec}", lambda: eng.score(xs), eng.flops_per_image(1000) * bb, a.seconds)
        eng.close()
        del eng, xs
    sys.exit(0)

b = a.batch
cfg = synth.CFGS["ViT-B/16"]
eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=b)
eng.set_text_bank(synth.synth_unit_bank(1000, cfg.proj, 3), already_unit=True)
g = torch.Generator(device="cuda").manual_seed(0)
S, D, F, H = 197, 768, 3072, 12
M = b * S
x = torch.randn(b, 3, 224, 224, device="cuda", generator=g)
act = torch.randn(M, D, device="cuda", generator=g).to(torch.float16)
hid = torch.randn(M, F, device="cuda", generator=g).to(torch.float16)
w_qkv = (torch.randn(3 * D, D, device="cuda", generator=g) * D ** -0.5).to(torch.float16)
w_fc1 = (torch.randn(F, D, device="cuda", generator=g) * D ** -0.5).to(torch.float16)
w_fc2 = (torch.randn(D, F, device="cuda", generator=g) * F ** -0.5).to(torch.float16)
w_out = (torch.randn(D, D, device="cuda", generator=g) * D ** -0.5).to(torch.float16)
stats = torch.randn(2 * (D // 256), M, 2, device="cuda", generator=g).abs() + 1.0
stats[..., 1] += stats[..., 0] ** 2
bias_qkv, c_qkv = torch.randn(3 * D, device="cuda", generator=g), torch.randn(3 * D, device="cuda", generator=g)
bias_fc1, c_fc1 = torch.randn(F, device="cuda", generator=g), torch.randn(F, device="cuda", generator=g)
bias_d = torch.randn(D, device="cuda", generator=g) * 0.01
x_hi = torch.randn(M, D, device="cuda", generator=g).to(torch.float16)
x_lo = torch.zeros(M, D, device="cuda", dtype=torch.float16)
qkv = (torch.randn(M, 3 * D, device="cuda", generator=g) * 0.5).to(torch.float16)
big = torch.randn(8192, 8192, device="cuda", generator=g).to(torch.bfloat16)
big2 = torch.randn(8192, 8192, device="cuda", generator=g).to(torch.bfloat16)
w_fc2_small, w_out_small = w_fc2 * 0.01, w_out * 0.01   # keep the in-place residual loops from growing
big_h, big2_h = big.to(torch.float16), big2.to(torch.float16)
zero_b = torch.zeros(8192, 8192, device="cuda", dtype=torch.bfloat16)
zero_h = torch.zeros(8192, 8192, device="cuda", dtype=torch.float16)
act_zero = torch.zeros_like(act)
cp_src = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
cp_dst = torch.empty_like(cp_src)

WORK = [
    ("step (ViT-B/16 forward + MCM tail)", lambda: eng.score(x), eng.flops_per_image(1000) * b),
    ("gemm q/k/v 2304x768 (LN fold, fp16 out)", lambda: eng.dbg_gemm_ln(act, w_qkv, bias_qkv, c_qkv, stats, D), 2.0 * M * 3 * D * D),
    ("gemm fc1 3072x768 (LN fold, quick_gelu, fp16 out)", lambda: eng.dbg_gemm_ln(act, w_fc1, bias_fc1, c_fc1, stats, D, gelu=True), 2.0 * M * F * D),
    ("gemm fc2 768x3072 (residual pair)", lambda: eng.dbg_gemm_resid_h2(hid, w_fc2_small, bias_d, x_hi, x_lo), 2.0 * M * D * F),
    ("gemm out_proj 768x768 (residual pair)", lambda: eng.dbg_gemm_resid_h2(act, w_out_small, bias_d, x_hi, x_lo), 2.0 * M * D * D),
    ("attention (512 x 12 heads, 197 tokens)", lambda: eng.dbg_attention(qkv, b, S, H), 4.0 * b * H * S * S * 64),
    ("cuBLAS bf16 8192^3 (torch.matmul)", lambda: torch.matmul(big, big2), 2.0 * 8192 ** 3),
    ("cuBLAS fp16 q/k/v shape, no epilogue (torch.matmul)", lambda: torch.matmul(act, w_qkv.t()), 2.0 * M * 3 * D * D),
    ("cuBLAS fp16 fc2 shape, no epilogue (torch.matmul)", lambda: torch.matmul(hid, w_fc2.t()), 2.0 * M * D * F),
    # operand format and operand data: what the tensor pipe's energy depends on
    ("cuBLAS fp16 8192^3 (torch.matmul)", lambda: torch.matmul(big_h, big2_h), 2.0 * 8192 ** 3),
    ("cuBLAS bf16 8192^3, all-zero operands", lambda: torch.matmul(zero_b, zero_b), 2.0 * 8192 ** 3),
    ("cuBLAS fp16 8192^3, all-zero operands", lambda: torch.matmul(zero_h, zero_h), 2.0 * 8192 ** 3),
    ("gemm q/k/v 2304x768, all-zero A", lambda: eng.dbg_gemm_ln(act_zero, w_qkv, bias_qkv, c_qkv, stats, D), 2.0 * M * 3 * D * D),
    # 1 GiB read + 1 GiB written per call: what a DRAM byte costs ("flops" = bytes here: pj_per_flop reads as pJ per byte)
    ("copy 1 GiB (dst.copy_(src); pj_per_flop = pJ per DRAM byte)", lambda: cp_dst.copy_(cp_src), 2.0 * (1 << 30)),
]
only = [s for s in a.only.split(",") if s]
for name, fn, fl in WORK:
    if only and not any(o in name for o in only):
        continue
    run_phase(name, fn, fl, a.seconds)
