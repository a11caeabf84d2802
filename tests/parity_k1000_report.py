"""Stream-level parity at the configs the headline metric is quoted on (BASELINE configs[2..4]):
ViT-B/16 or ViT-L/14, K = 1000 prototype bank, n_id ID + n_ood OOD images, engine vs the fp32 oracle
restatement run on the same GPU (TF32 off).  Prints one JSON line per (config, noise, precision) with the
score error, AUROC / FPR95 of both sides and their differences, and keeps the raw score vectors under
gpurun_out/ for offline analysis.  The same harness is what tests/test_gpu_parity_k1000.py asserts on.

    python tests/parity_k1000_report.py --cfg ViT-B/16 --n-id 5000 --n-ood 10000 --noise 0.8 --precision 0 1
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(cfg_name, K, n_id, n_ood, noise, precisions, batch, out_dir, wseed=5, tag=""):
    from mcm_b200 import synth
    from mcm_b200.engine import McmEngine
    from oracle import clip_mcm_oracle as O
    from k1000_harness import K1000Harness
    h = K1000Harness(cfg_name, K=K, noise=noise, wseed=wseed)
    t0 = time.time()
    ref_id, ref_ood = h.oracle_scores(n_id, n_ood)
    t_ref = time.time() - t0
    m_ref = O.get_measures(-ref_id, -ref_ood)
    spread = float(np.concatenate([ref_id, ref_ood]).std())
    eng = McmEngine.from_state_dict(h.sd, h.cfg, max_batch=batch)
    try:
        eng.set_text_bank(h.bank)
        for prec in precisions:
            if prec:
                eng.set_precision(prec)
            t0 = time.time()
            got_id, got_ood = h.engine_scores(eng, n_id, n_ood, batch)
            t_eng = time.time() - t0
            m_got = O.get_measures(-got_id, -got_ood)
            err = np.concatenate([got_id - ref_id, got_ood - ref_ood])
            line = dict(cfg=cfg_name, K=K, n_id=n_id, n_ood=n_ood, noise=noise, precision=prec,
                        score_mean=float(ref_id.mean()), score_std=spread,
                        max_abs_err=float(np.abs(err).max()), rms_err=float(np.sqrt((err.astype(np.float64) ** 2).mean())),
                        max_err_over_std=float(np.abs(err).max() / spread),
                        auroc_ref=float(m_ref[0]), auroc=float(m_got[0]), d_auroc=float(abs(m_got[0] - m_ref[0])),
                        fpr_ref=float(m_ref[2]), fpr=float(m_got[2]), d_fpr=float(abs(m_got[2] - m_ref[2])),
                        t_oracle_s=round(t_ref, 1), t_engine_s=round(t_eng, 1))
            print(json.dumps(line), flush=True)
            if out_dir:
                np.savez_compressed(os.path.join(out_dir, f"k1000_{cfg_name.replace('/', '')}_noise{noise}_p{prec}{tag}.npz"),
                                    got_id=got_id, got_ood=got_ood, ref_id=ref_id, ref_ood=ref_ood)
    finally:
        eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="ViT-B/16")
    ap.add_argument("--K", type=int, default=1000)
    ap.add_argument("--n-id", type=int, default=5000)
    ap.add_argument("--n-ood", type=int, default=10000)
    ap.add_argument("--noise", type=float, nargs="+", default=[0.8])
    ap.add_argument("--precision", type=int, nargs="+", default=[0])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for nz in a.noise:
        run(a.cfg, a.K, a.n_id, a.n_ood, nz, a.precision, a.batch, a.out)


if __name__ == "__main__":
    main()
