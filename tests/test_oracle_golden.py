"""The CPU oracle restatement against the golden vectors produced by the UNMODIFIED reference
(utils/detection_util.py:209-249 + HF CLIP) in the authoring container -- the pin that lets the
oracle be trusted as the checker of the CUDA path."""
import os

import numpy as np
import pytest
import torch

from helpers import golden_inputs
from oracle import clip_mcm_oracle as O


@pytest.mark.parametrize("case", ["tiny_text_k10", "tiny_proto_k16_T2", "small_proto_k20"])
def test_oracle_reproduces_reference_scores(case, golden_dir):
    z = np.load(os.path.join(golden_dir, case + ".npz"))
    cfg, sd, _protos, id_imgs, ood_imgs = golden_inputs(z)
    torch.set_num_threads(os.cpu_count() or 1)
    for sc in [str(s) for s in z["scores"]]:
        key = sc.replace("-", "_")
        o_in = O.ood_scores(id_imgs, sd, cfg, z["bank"], T=int(z["T"]), score=sc, batch=64)
        o_out = O.ood_scores(ood_imgs, sd, cfg, z["bank"], T=int(z["T"]), score=sc, batch=64)
        assert o_in.dtype == np.float32
        scale = max(1.0, float(np.abs(z[f"ref_in_{key}"]).max()))
        np.testing.assert_allclose(o_in, z[f"ref_in_{key}"], rtol=0, atol=4e-6 * scale)
        np.testing.assert_allclose(o_out, z[f"ref_out_{key}"], rtol=0, atol=4e-6 * scale)
        m = O.get_measures(-z[f"ref_in_{key}"], -z[f"ref_out_{key}"])
        np.testing.assert_allclose(m, z[f"measures_{key}"], rtol=0, atol=1e-12)


def test_oracle_b16_subset(golden_dir):
    """ViT-B/16 (BASELINE config 1): 12 of the 256 ID images, to keep the CPU suite short."""
    z = np.load(os.path.join(golden_dir, "b16_text_k10_cfg1.npz"))
    cfg, sd, _protos, id_imgs, _ood = golden_inputs(dict(z, n_id=12, n_ood=1))
    torch.set_num_threads(os.cpu_count() or 1)
    o_in = O.ood_scores(id_imgs[:12], sd, cfg, z["bank"], T=1, score="MCM", batch=12)
    np.testing.assert_allclose(o_in, z["ref_in_MCM"][:12], rtol=0, atol=4e-6)


def test_flops_formula():
    assert O.flops_per_image(O.CFGS["ViT-B/16"], 1000) == pytest.approx(35.128e9, rel=2e-4)
    assert O.flops_per_image(O.CFGS["ViT-L/14"], 1000) == pytest.approx(162.027e9, rel=2e-4)


def test_oracle_preprocess_matches_torchvision():
    """oracle.preprocess_u8 == the ToTensor -> Normalize tail of the reference val_preprocess
    (utils/train_eval_util.py:27-34) executed by torchvision itself on PIL images, bit for bit."""
    PIL = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    from mcm_b200 import synth
    u8 = synth.synth_images_u8(3, 11)
    tail = T.Compose([T.ToTensor(), T.Normalize(mean=O.CLIP_MEAN, std=O.CLIP_STD)])
    ref = torch.stack([tail(PIL.fromarray(u8[i])) for i in range(u8.shape[0])])
    got = O.preprocess_u8(u8)
    assert got.dtype == torch.float32 and got.shape == (3, 3, 224, 224)
    assert torch.equal(got, ref)


RESIZE_SIZES = [(375, 500), (500, 375), (224, 224), (224, 300), (301, 224), (225, 224), (100, 160), (160, 100),
                (333, 1000), (1500, 431), (64, 64), (227, 229), (900, 1200)]


def test_resize_oracle_matches_torchvision():
    """oracle.pil_resize_oracle == CenterCrop(224)(Resize(224)(pil_image)) executed by torchvision + Pillow themselves
    (the first two steps of the reference val_preprocess, utils/train_eval_util.py:29-31), bit for bit, over
    down-scaling, up-scaling, both orientations, odd crop offsets and the no-op size."""
    PIL = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    from oracle import pil_resize_oracle as R
    tf = T.Compose([T.Resize(224), T.CenterCrop(224)])
    rng = np.random.default_rng(3)
    for h, w in RESIZE_SIZES:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref = np.asarray(tf(PIL.fromarray(img)))
        got = R.resize_center_crop_u8(img)
        assert got.shape == (224, 224, 3) and got.dtype == np.uint8
        np.testing.assert_array_equal(got, ref, err_msg=f"{h}x{w}")


def test_resize_planner_tables_match_oracle():
    """The host half of mcm_resize_crop_u8 (C++, csrc/resize.cuh: size / crop rules + Pillow's coefficient tables in
    fixed point) against the oracle restatement -- runs without a GPU through mcm_dbg_resize_tables."""
    import ctypes as C
    from mcm_b200 import _lib
    from oracle import pil_resize_oracle as R
    lib = _lib.load()
    for h, w in RESIZE_SIZES + [(3000, 4000), (4000, 3000), (1, 7)]:
        cap = 224 * (2 + 600)
        th, tv, ks = np.zeros(cap, np.int32), np.zeros(cap, np.int32), (C.c_int32 * 2)()
        assert lib.mcm_dbg_resize_tables(h, w, 224, ks, C.c_void_p(th.ctypes.data), C.c_void_p(tv.ctypes.data), cap) == 0
        th = th[:224 * (2 + ks[0])].reshape(224, -1)
        tv = tv[:224 * (2 + ks[1])].reshape(224, -1)
        nh, nw = R.resized_size(h, w)
        top, left = R.crop_offsets(nh, nw)
        bh, ch = R.precompute_coeffs(w, nw)
        bv, cv = R.precompute_coeffs(h, nh)
        np.testing.assert_array_equal(th[:, :2], bh[left:left + 224], err_msg=f"{h}x{w} bounds_h")
        np.testing.assert_array_equal(th[:, 2:], ch[left:left + 224], err_msg=f"{h}x{w} kk_h")
        np.testing.assert_array_equal(tv[:, :2], bv[top:top + 224], err_msg=f"{h}x{w} bounds_v")
        np.testing.assert_array_equal(tv[:, 2:], cv[top:top + 224], err_msg=f"{h}x{w} kk_v")
    # a down-scaling factor beyond the shared-memory tile budget is refused, not mis-computed
    assert lib.mcm_dbg_resize_tables(20000, 20000, 224, ks, C.c_void_p(th.ctypes.data), C.c_void_p(tv.ctypes.data), cap) == _lib.EUNSUPPORTED


def test_resize_oracle_matches_golden_digests(golden_dir):
    """Same pin without Pillow / torchvision: SHA-1 digests of their outputs on seeded images, made here by
    oracle/make_golden_resize.py."""
    import hashlib
    from oracle import pil_resize_oracle as R
    from oracle.make_golden_resize import image
    z = np.load(os.path.join(golden_dir, "resize_crop_pil.npz"))
    for i, ((h, w), want) in enumerate(zip(z["sizes"], z["sha1"])):
        got = R.resize_center_crop_u8(image(int(h), int(w), i))
        assert hashlib.sha1(got.tobytes()).hexdigest() == str(want), f"{h}x{w}"


@pytest.mark.parametrize("tag,normalize", [("u", False), ("n", True)])
def test_maha_oracle_reproduces_reference(golden_dir, tag, normalize):
    """Mahalanobis baseline: the oracle restatement against outputs of the UNMODIFIED reference get_mean_prec /
    get_Mahalanobis_score (utils/detection_util.py:148-207) stored by oracle/make_golden_maha.py -- statistics,
    scores, and the dropped trailing OOD batch."""
    from oracle.make_golden_maha import SPEC, build_inputs
    z = np.load(os.path.join(golden_dir, "maha_tiny.npz"))
    cfg, sd, train, train_labels, id_imgs, ood = build_inputs()
    B = SPEC["batch"]
    with torch.no_grad():
        feats = [O.image_features(torch.from_numpy(train[s:s + B]), sd, cfg) for s in range(0, len(train), B)]
    labs = [train_labels[s:s + B] for s in range(0, len(train), B)]
    mean, prec = O.maha_mean_prec(feats, labs, SPEC["n_cls"], normalize)
    np.testing.assert_allclose(mean.numpy(), z[f"mean_{tag}"], rtol=1e-4, atol=1e-6)
    ref_mean, ref_prec = torch.from_numpy(z[f"mean_{tag}"]), torch.from_numpy(z[f"prec_{tag}"])
    o_in = O.maha_scores(id_imgs, sd, cfg, ref_mean, ref_prec, B, normalize, True)
    o_out = O.maha_scores(ood, sd, cfg, ref_mean, ref_prec, B, normalize, False)
    assert o_out.shape == z[f"ref_out_{tag}"].shape == ((len(ood) // B) * B,)
    scale = float(np.abs(z[f"ref_in_{tag}"]).max())
    np.testing.assert_allclose(o_in, z[f"ref_in_{tag}"], rtol=0, atol=2e-4 * scale)
    np.testing.assert_allclose(o_out, z[f"ref_out_{tag}"], rtol=0, atol=2e-4 * scale)
