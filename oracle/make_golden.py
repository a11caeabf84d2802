"""Generate the golden fixtures under tests/golden/ -- TEST INFRASTRUCTURE.

Runs ONLY in the authoring container (needs /root/reference).  For every case it
  1. builds seeded synthetic weights / images (mcm_b200.synth, numpy Philox -> the
     same bytes can be regenerated on the GPU box from the seeds stored here),
  2. runs the reference's own, unmodified ``get_ood_scores_clip`` +
     ``get_measures`` on CPU through oracle/reference_shims.py,
  3. runs the torch restatement (oracle/clip_mcm_oracle.py) on the same inputs and
     asserts it agrees with the reference (scores to 2e-6, metrics exactly),
  4. stores seeds, bank, reference scores and reference metrics in a small .npz.

Usage:  python oracle/make_golden.py [case ...]      (default: all cases)
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcm_b200 import synth  # noqa: E402
from oracle import clip_mcm_oracle as O  # noqa: E402
from oracle import reference_shims as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

IMAGENET10 = ["brambling bird", "bull frog", "Siamese cat", "swiss mountain dog", "horse", "antelope",
              "container ship", "garbage truck", "sports car", "warplane"]  # utils/common.py:36-46 (sorted by wnid)

# name -> spec.  kind "text": bank = the (random-init) text tower of the shim model on
# fake-tokenised "a photo of a {c}" prompts, captured from inside the reference loop.
# kind "proto": bank = centred prototype bank (mcm_b200.synth.centred_prototype_bank),
# returned by the shim's get_text_features.
CASES = {
    "tiny_text_k10": dict(cfg="tiny", kind="text", K=10, n_id=96, n_ood=80, batch=32, wseed=5,
                          scores=["MCM", "energy", "max-logit", "entropy", "var"], T=1),
    "tiny_proto_k16_T2": dict(cfg="tiny", kind="proto", K=16, n_id=128, n_ood=128, batch=48, wseed=7, noise=0.6,
                              scores=["MCM"], T=2),
    "small_proto_k20": dict(cfg="small", kind="proto", K=20, n_id=256, n_ood=256, batch=64, wseed=5, noise=0.6,
                            scores=["MCM", "energy", "max-logit", "entropy", "var"], T=1),
    "b16_text_k10_cfg1": dict(cfg="ViT-B/16", kind="text", K=10, n_id=256, n_ood=128, batch=256, wseed=5,
                              scores=["MCM"], T=1),
    "b16_proto_k100": dict(cfg="ViT-B/16", kind="proto", K=100, n_id=1024, n_ood=1024, batch=128, wseed=5,
                           noise=0.8, scores=["MCM"], T=1),
}


def build_inputs(spec):
    """Everything here is regenerated identically by tests/ on the GPU box."""
    cfg = synth.CFGS[spec["cfg"]]
    sd = synth.synth_vision_state_dict(cfg, spec["wseed"])
    if spec["kind"] == "proto":
        protos = synth.synth_images(spec["K"], 100)
        id_imgs = synth.synth_prototype_stream(spec["n_id"], protos, 1, spec["noise"])
        ood = synth.synth_images(spec["n_ood"], 2, std=float(np.sqrt(1.0 + spec["noise"] ** 2)))
    else:
        protos = None
        id_imgs = synth.synth_images(spec["n_id"], 1)
        ood = synth.synth_images(spec["n_ood"], 2, mean=0.3, std=1.5)
    return cfg, sd, protos, id_imgs, ood


def make_case(name, spec):
    t0 = time.time()
    cfg, sd, protos, id_imgs, ood = build_inputs(spec)
    model = R.make_shim_clip(cfg, sd, seed=spec["wseed"])
    labels = IMAGENET10[: spec["K"]] if spec["K"] <= 10 else [f"class {i}" for i in range(spec["K"])]
    if spec["kind"] == "proto":
        with torch.no_grad():
            pf = torch.cat([O.image_features(torch.from_numpy(protos[i:i + 32]), sd, cfg)
                            for i in range(0, spec["K"], 32)]).numpy()
        bank = synth.centred_prototype_bank(pf)
        type(model).bank_override = torch.from_numpy(bank)
    else:
        type(model).bank_override = None
    out = dict(cfg=spec["cfg"], kind=spec["kind"], K=spec["K"], n_id=spec["n_id"], n_ood=spec["n_ood"],
               wseed=spec["wseed"], noise=spec.get("noise", 0.0), T=spec["T"], scores=np.array(spec["scores"]))
    for sc in spec["scores"]:
        ref_in = R.run_reference_scores(model, id_imgs, labels, T=spec["T"], score=sc, batch_size=spec["batch"])
        ref_out = R.run_reference_scores(model, ood, labels, T=spec["T"], score=sc, batch_size=spec["batch"])
        if spec["kind"] == "text":
            cb = type(model).captured_bank
            bank_t = cb / cb.norm(dim=-1, keepdim=True)   # utils/detection_util.py:231
            bank = bank_t.numpy()
        else:
            # the reference re-normalises the (already unit) rows, :231 -- do the same
            bt = torch.from_numpy(bank)
            bank = (bt / bt.norm(dim=-1, keepdim=True)).numpy()
        o_in = O.ood_scores(id_imgs, sd, cfg, bank, T=spec["T"], score=sc, batch=spec["batch"])
        o_out = O.ood_scores(ood, sd, cfg, bank, T=spec["T"], score=sc, batch=spec["batch"])
        err = max(np.abs(ref_in - o_in).max(), np.abs(ref_out - o_out).max())
        scale = max(np.abs(ref_in).max(), 1e-30)
        assert ref_in.dtype == np.float32 and ref_in.shape == (spec["n_id"],)
        assert err <= 2e-6 * max(1.0, scale), (name, sc, err)
        m_ref = R.run_reference_measures(ref_in, ref_out)
        m_orc = O.get_measures(-ref_in, -ref_out)
        assert np.allclose(m_ref, m_orc, rtol=0, atol=1e-12), (m_ref, m_orc)
        key = sc.replace("-", "_")
        out[f"ref_in_{key}"] = ref_in
        out[f"ref_out_{key}"] = ref_out
        out[f"measures_{key}"] = np.asarray(m_ref, dtype=np.float64)
        print(f"[{name}] {sc}: oracle-vs-reference max|d|={err:.3e}  in mean/std {ref_in.mean():.6g}/{ref_in.std():.3g} "
              f"out {ref_out.mean():.6g}/{ref_out.std():.3g}  AUROC/AUPR/FPR95={m_ref}", flush=True)
    out["bank"] = bank.astype(np.float32)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"[{name}] written in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    assert R.REFERENCE_AVAILABLE, "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(CASES)
    for n in names:
        make_case(n, CASES[n])
