"""GPU-box diagnostic for the tcgen05 GEMM: isolates descriptor / pipeline mistakes.

Each case runs in its own process (a device-side trap poisons the CUDA context) and appends one
JSON line to gpurun_out/gemm_diag.jsonl: max error, mismatch fraction and where the mismatches sit
(by row mod 8, by 8-column group, by which 16-wide K slice of A is non-zero).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "gemm_diag.jsonl")

CASES = [
    # (M, N, K, epi, kslice)  kslice: only A[:, 16*kslice : 16*kslice+16] non-zero (-1 = all)
    (128, 128, 64, 2, 0), (128, 128, 64, 2, 1), (128, 128, 64, 2, 3), (128, 128, 64, 2, -1),
    (128, 128, 128, 2, -1), (128, 256, 64, 2, -1), (256, 128, 64, 2, -1), (128, 128, 64, 0, -1),
    (128, 128, 1024, 2, -1), (512, 512, 512, 2, -1), (197 * 64, 768, 768, 2, -1), (197 * 256, 3072, 768, 1, -1),
]


def run_case(i):
    import torch
    from mcm_b200 import synth
    from mcm_b200.engine import McmEngine
    M, N, K, epi, ks = CASES[i]
    cfg = synth.CFGS["tiny"]
    eng = McmEngine.from_state_dict(synth.synth_vision_state_dict(cfg, 5), cfg, max_batch=4)
    g = torch.Generator(device="cuda").manual_seed(i)
    a = torch.randn(M, K, device="cuda", generator=g)
    if ks >= 0:
        m = torch.zeros(K, device="cuda")
        m[16 * ks:16 * ks + 16] = 1
        a = a * m
    a = a.to(torch.float16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.float16)
    bias = torch.zeros(N, device="cuda")
    resid = torch.zeros(M, N, device="cuda") if epi == 2 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = eng.dbg_gemm(a, w, bias, resid, epi)
    torch.cuda.synchronize()
    ev0.record()
    reps = 5
    for _ in range(reps):
        out = eng.dbg_gemm(a, w, bias, resid, epi)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    ref = a.float() @ w.float().t()
    if epi == 1:
        ref = ref * torch.sigmoid(1.702 * ref)
    d = (out.float() - ref).abs()
    tol = 0.05 if epi != 2 else 2e-3
    bad = d > tol
    rec = dict(case=i, M=M, N=N, K=K, epi=epi, kslice=ks, max_err=float(d.max()), bad_frac=float(bad.float().mean()),
               ms=ms, tflops=2.0 * M * N * K / ms / 1e9, ref_absmax=float(ref.abs().max()), out_absmax=float(out.float().abs().max()),
               nan=bool(torch.isnan(out.float()).any()))
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        rec["bad_rows_mod8"] = sorted(set((rows % 8).tolist()))
        rec["bad_rows_first"] = rows[:8].tolist()
        rec["bad_cols_first"] = cols[:8].tolist()
        rec["bad_row_blocks128"] = sorted(set((rows // 128).tolist()))[:8]
        rec["bad_col_blocks32"] = sorted(set((cols // 32).tolist()))[:16]
        rec["sample_out"] = out.float()[int(rows[0]), :4].tolist()
        rec["sample_ref"] = ref[int(rows[0]), :4].tolist()
    print(json.dumps(rec))
    with open(OUT, "a") as f:
        f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        for i in range(len(CASES)):
            r = subprocess.run([sys.executable, __file__, str(i)], capture_output=True, text=True, timeout=300)
            if r.returncode != 0:
                with open(OUT, "a") as f:
                    f.write(json.dumps(dict(case=i, spec=CASES[i], rc=r.returncode, stderr=r.stderr[-1500:], stdout=r.stdout[-500:])) + "\n")
                print("case", i, "failed rc", r.returncode, r.stderr[-800:])
            else:
                print(r.stdout.strip()[-600:])
