"""Golden fixture for the Mahalanobis baseline (``--score maha``) -- TEST INFRASTRUCTURE.

Runs the UNMODIFIED ``get_mean_prec`` and ``get_Mahalanobis_score`` of the reference
(``utils/detection_util.py:148-207``) on CPU through oracle/reference_shims.py (a tiny HF CLIP, ``.cuda()`` no-op),
checks the restatement in oracle/clip_mcm_oracle.py against them and stores seeds + reference outputs in
``tests/golden/maha_tiny.npz``.     python -m oracle.make_golden_maha
"""
import os
import tempfile
import types

import numpy as np
import torch

from mcm_b200 import synth
from oracle import clip_mcm_oracle as O
from oracle import reference_shims as R

SPEC = dict(cfg="tiny", wseed=5, n_cls=6, n_train=96, n_id=50, n_ood=50, batch=16, noise=0.6)


def build_inputs(spec=SPEC):
    cfg = synth.CFGS[spec["cfg"]]
    sd = synth.synth_vision_state_dict(cfg, spec["wseed"])
    protos = synth.synth_images(spec["n_cls"], 100)
    rng = np.random.default_rng(77)
    train_labels = rng.integers(0, spec["n_cls"], size=spec["n_train"])
    train = synth.synth_images(spec["n_train"], 11) * np.float32(spec["noise"]) + protos[train_labels]
    id_imgs = synth.synth_prototype_stream(spec["n_id"], protos, 1, spec["noise"])
    ood = synth.synth_images(spec["n_ood"], 2, std=float(np.sqrt(1.0 + spec["noise"] ** 2)))
    return cfg, sd, train.astype(np.float32), train_labels.astype(np.int64), id_imgs, ood


class _LabelLoader(R._ListLoader):
    def __init__(self, images, labels, batch_size):
        super().__init__(images, batch_size)
        self.labels = torch.as_tensor(labels)

    def __iter__(self):
        for s in range(0, self.images.shape[0], self.batch_size):
            yield self.images[s:s + self.batch_size], self.labels[s:s + self.batch_size]


def main():
    assert R.REFERENCE_AVAILABLE, "needs /root/reference"
    du = R.load_reference_detection_util()
    cfg, sd, train, train_labels, id_imgs, ood = build_inputs()
    model = R.make_shim_clip(cfg, sd, seed=SPEC["wseed"])
    out = {k: v for k, v in SPEC.items()}
    for normalize in (False, True):
        with tempfile.TemporaryDirectory() as tmp, R.cpu_cuda_noop():
            args = types.SimpleNamespace(model="CLIP", n_cls=SPEC["n_cls"], feat_dim=cfg.proj, gpu="cpu", normalize=normalize,
                                         template_dir=tmp, in_dataset="synthetic", max_count=0, batch_size=SPEC["batch"])
            mean, prec = du.get_mean_prec(args, model, _LabelLoader(train, train_labels, SPEC["batch"]))
            ref_in = du.get_Mahalanobis_score(args, model, _LabelLoader(id_imgs, np.zeros(len(id_imgs), np.int64), SPEC["batch"]),
                                              mean, prec, in_dist=True)
            ref_out = du.get_Mahalanobis_score(args, model, _LabelLoader(ood, np.zeros(len(ood), np.int64), SPEC["batch"]),
                                               mean, prec, in_dist=False)
        # the restatement on the same inputs
        with torch.no_grad():
            feats = [O.image_features(torch.from_numpy(train[s:s + SPEC["batch"]]), sd, cfg) for s in range(0, len(train), SPEC["batch"])]
        labs = [train_labels[s:s + SPEC["batch"]] for s in range(0, len(train), SPEC["batch"])]
        o_mean, o_prec = O.maha_mean_prec(feats, labs, SPEC["n_cls"], normalize)
        assert torch.allclose(o_mean, mean.float(), rtol=1e-4, atol=1e-6), (o_mean - mean).abs().max()
        o_in = O.maha_scores(id_imgs, sd, cfg, mean, prec, SPEC["batch"], normalize, True)
        o_out = O.maha_scores(ood, sd, cfg, mean, prec, SPEC["batch"], normalize, False)
        assert ref_in.shape == o_in.shape and ref_out.shape == o_out.shape == ((len(ood) // SPEC["batch"]) * SPEC["batch"],)
        scale = float(np.abs(ref_in).max())
        err = max(np.abs(ref_in - o_in).max(), np.abs(ref_out - o_out).max()) / scale
        print(f"normalize={normalize}: cond(precision)={float(torch.linalg.cond(prec)):.3g}  scores in {ref_in.mean():.4g} out {ref_out.mean():.4g}"
              f"  oracle-vs-reference rel err {err:.2e}")
        assert err <= 2e-4, err
        tag = "n" if normalize else "u"
        out[f"mean_{tag}"], out[f"prec_{tag}"] = mean.numpy(), prec.numpy()
        out[f"ref_in_{tag}"], out[f"ref_out_{tag}"] = ref_in, ref_out
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maha_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
