"""Mahalanobis baseline (`--score maha`, utils/detection_util.py:148-207) on the B200 engine against the oracle
restatement (pinned to the unmodified reference by tests/golden/maha_tiny.npz)."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import report

pytestmark = pytest.mark.gpu


class LabelLoader:
    def __init__(self, images, labels, batch_size):
        self.images, self.labels, self.batch_size = torch.as_tensor(images), torch.as_tensor(labels), batch_size
        self.dataset = range(self.images.shape[0])

    def __len__(self):
        return -(-self.images.shape[0] // self.batch_size)

    def __iter__(self):
        for s in range(0, self.images.shape[0], self.batch_size):
            yield self.images[s:s + self.batch_size], self.labels[s:s + self.batch_size]


@pytest.mark.parametrize("normalize", [False, True])
def test_maha_tail_matches_quadratic_form(engine_factory, normalize):
    """Kernel level: given the SAME fp32 features, 0.5 min_k |f L - mu_k L|^2 (one whitening GEMM + squared distances)
    equals the reference's K-iteration quadratic forms."""
    from oracle import clip_mcm_oracle as O
    eng, _, cfg = engine_factory("tiny", 5, 32)
    g = torch.Generator().manual_seed(3)
    P, K, b = cfg.proj, 37, 29
    train = torch.randn(400, P, generator=g) * torch.linspace(0.2, 2.0, P) + torch.randn(P, generator=g)
    mean = torch.randn(K, P, generator=g)
    prec = torch.linalg.inv(torch.cov(train.T.double())).float()
    feats = torch.randn(b, P, generator=g) * 1.5
    eng.set_maha(mean, prec, normalize)
    got = eng.dbg_maha_from_features(feats.cuda()).cpu().numpy()
    ref = O.maha_scores_from_features(feats, mean, prec, normalize)
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-4)
    with pytest.raises(ValueError):
        eng.set_maha(mean, -prec, normalize)              # not positive definite
    with pytest.raises(ValueError):
        eng.set_maha(mean[:, :-1], prec, normalize)


@pytest.mark.parametrize("tag,normalize", [("u", False), ("n", True)])
def test_maha_end_to_end(engine_factory, golden_dir, tag, normalize):
    """Drop-in functions: get_mean_prec + get_Mahalanobis_score with reference signatures.  The statistics are compared
    with the reference's; the scores with the reference's scores for the SAME statistics.  Mahalanobis distances
    amplify feature error by the conditioning of the precision matrix (cond ~2e3 here), so the bound on the scores is
    relative: 1 % of the score scale (measured 0.2 %) -- far below the ID / OOD gap -- and the AUROC must agree."""
    from mcm_b200 import detection_util as DU
    from mcm_b200.engine import B200ClipNet
    from oracle import clip_mcm_oracle as O
    from oracle.make_golden_maha import SPEC, build_inputs
    z = np.load(os.path.join(golden_dir, "maha_tiny.npz"))
    eng, sd, cfg = engine_factory("tiny", 5, 32)
    net = B200ClipNet(eng).eval()
    _cfg, _sd, train, train_labels, id_imgs, ood = build_inputs()
    B = SPEC["batch"]
    args = types.SimpleNamespace(model="CLIP", n_cls=SPEC["n_cls"], feat_dim=cfg.proj, normalize=normalize, template_dir=None,
                                 in_dataset="synthetic", max_count=0, batch_size=B)
    mean, prec = DU.get_mean_prec(args, net, LabelLoader(train, train_labels, B))
    np.testing.assert_allclose(mean.numpy(), z[f"mean_{tag}"], rtol=0, atol=5e-3 * float(np.abs(z[f"mean_{tag}"]).max()))
    ref_mean, ref_prec = torch.from_numpy(z[f"mean_{tag}"]), torch.from_numpy(z[f"prec_{tag}"])
    got_in = DU.get_Mahalanobis_score(args, net, LabelLoader(id_imgs, np.zeros(len(id_imgs), np.int64), B), ref_mean, ref_prec, in_dist=True)
    got_out = DU.get_Mahalanobis_score(args, net, LabelLoader(ood, np.zeros(len(ood), np.int64), B), ref_mean, ref_prec, in_dist=False)
    ref_in, ref_out = z[f"ref_in_{tag}"], z[f"ref_out_{tag}"]
    assert got_in.dtype == np.float32 and got_in.shape == ref_in.shape and got_out.shape == ref_out.shape
    scale = float(np.abs(ref_in).max())
    err = max(np.abs(got_in - ref_in).max(), np.abs(got_out - ref_out).max()) / scale
    m_got = DU.get_measures(-got_in, -got_out)
    m_ref = O.get_measures(-ref_in, -ref_out)
    report("maha", dict(normalize=normalize, rel_err=float(err), auroc=float(m_got[0]), auroc_ref=float(m_ref[0])))
    assert err <= 1e-2, err
    assert abs(m_got[0] - m_ref[0]) <= 5e-3, (m_got, m_ref)
