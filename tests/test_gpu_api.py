"""Drop-in boundary behaviour: reference signature / return contract / error behaviour
(utils/detection_util.py:209-249; HF modeling_clip.py:204-207), ragged and empty streams,
the host-stream C-ABI entry point, launch accounting."""
import numpy as np
import pytest
import torch

from helpers import ListLoader, make_args

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tiny_net(engine_factory):
    from mcm_b200 import synth
    from mcm_b200.engine import B200ClipNet
    eng, sd, cfg = engine_factory("tiny", 5, 32)
    bank = synth.synth_unit_bank(12, cfg.proj, 9)
    eng.set_text_bank(bank)          # tests that call the engine directly must not depend on test order
    return B200ClipNet(eng, text_bank=bank).eval(), cfg, bank


def test_return_contract_and_ragged_batches(tiny_net):
    from mcm_b200 import detection_util as DU
    from mcm_b200 import synth
    net, cfg, bank = tiny_net
    imgs = synth.synth_images(77, 3)
    labels = [f"c{i}" for i in range(12)]
    a = DU.get_ood_scores_clip(make_args(), net, ListLoader(imgs, 32), labels, in_dist=True)   # 32+32+13
    b = DU.get_ood_scores_clip(make_args(), net, ListLoader(imgs, 7), labels)                  # 11 ragged batches
    c = DU.get_ood_scores_clip(make_args(), net, ListLoader(imgs, 100), labels)                # loader batch > max_batch
    assert a.dtype == np.float32 and a.shape == (77,) and a.flags["OWNDATA"]
    np.testing.assert_array_equal(a, b)          # a score depends only on its own image: batch-invariant, bit-exact
    np.testing.assert_array_equal(a, c)
    assert np.all(a < 0) and np.all(a >= -1)     # -max softmax
    e = DU.get_ood_scores_clip(make_args(), net, ListLoader(imgs[:0], 8), labels)
    assert e.shape == (0,) and e.dtype == np.float32


def test_scores_are_deterministic(tiny_net):
    from mcm_b200 import synth
    net, cfg, _ = tiny_net
    x = torch.from_numpy(synth.synth_images(16, 4)).cuda()
    s1 = net.engine.score(x).clone()
    s2 = net.engine.score(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(s1, s2)


def test_host_stream_matches_device_path(tiny_net):
    from mcm_b200 import synth
    net, cfg, _ = tiny_net
    eng = net.engine
    imgs = synth.synth_images(70, 5)
    dev = torch.cat([eng.score(torch.from_numpy(imgs[s:s + 32]).cuda()) for s in range(0, 70, 32)]).cpu().numpy()
    pinned = torch.from_numpy(imgs).pin_memory()
    for batch in (32, 9):
        host = eng.score_stream_host(pinned, batch=batch)
        np.testing.assert_array_equal(host, dev)
    np.testing.assert_array_equal(eng.score_stream_host(imgs, batch=16), dev)    # pageable host memory works too
    assert eng.score_stream_host(imgs[:0]).shape == (0,)


def test_all_score_kinds_via_args(tiny_net):
    from mcm_b200 import detection_util as DU
    from mcm_b200 import synth
    from oracle import clip_mcm_oracle as O
    net, cfg, bank = tiny_net
    imgs = synth.synth_images(20, 6)
    labels = [f"c{i}" for i in range(12)]
    feats = net.get_image_features(pixel_values=torch.from_numpy(imgs).cuda()).cpu()
    for sc in ("MCM", "max-logit", "energy", "entropy", "var"):
        got = DU.get_ood_scores_clip(make_args(T=2, score=sc), net, ListLoader(imgs, 8), labels)
        ref = O.scores_from_features(feats, torch.from_numpy(bank), 2, sc)
        np.testing.assert_allclose(got, ref, rtol=3e-4, atol=3e-6, err_msg=sc)


def test_error_behaviour(tiny_net):
    from mcm_b200 import detection_util as DU
    from mcm_b200 import synth
    net, cfg, _ = tiny_net
    eng = net.engine
    with pytest.raises(ValueError, match="doesn't match model"):      # HF:204-207
        eng.score(torch.zeros(2, 3, 192, 192, device="cuda"))
    with pytest.raises(ValueError):
        eng.score(torch.zeros(2, 3, 224, 224))                          # host tensor on the device path
    with pytest.raises(ValueError):
        eng.score(torch.zeros(2, 3, 224, 224, device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        eng.score(torch.zeros(33, 3, 224, 224, device="cuda"))          # > max_batch
    with pytest.raises(ValueError):
        eng.score(torch.zeros(2, 3, 224, 224, device="cuda"), score="maha")
    with pytest.raises(ValueError):
        eng.score(torch.zeros(2, 3, 224, 224, device="cuda"), T=0)
    with pytest.raises(ValueError):
        eng.set_text_bank(np.zeros((4, cfg.proj + 1), np.float32))
    with pytest.raises(TypeError):
        DU.get_ood_scores_clip(make_args(), torch.nn.Linear(2, 2), ListLoader(synth.synth_images(1, 1), 1), ["a"])


def test_bank_and_weights_required(engine_factory):
    from mcm_b200 import synth
    from mcm_b200.engine import McmEngine
    cfg = synth.CFGS["tiny"]
    eng = McmEngine(cfg, max_batch=4)
    try:
        x = torch.zeros(1, 3, 224, 224, device="cuda")
        with pytest.raises(RuntimeError, match="not finalized"):
            eng.score(x)
        sd = synth.synth_vision_state_dict(cfg, 5)
        partial = {k: v for k, v in sd.items() if "layers.1.mlp.fc2.weight" not in k}
        with pytest.raises(RuntimeError, match="never loaded"):
            eng.load_state_dict(partial)
        used = eng.load_state_dict({**sd, "logit_scale": torch.tensor(1.0), "text_model.foo": torch.zeros(3)})
        assert len(used) == len(sd)
        with pytest.raises(RuntimeError, match="bank is not set"):
            eng.score(x)
        eng.image_features(x)       # needs no bank
        bad = dict(sd)
        bad["visual_projection.weight"] = torch.zeros(3, 3)
        with pytest.raises(ValueError, match="expected"):
            eng.load_state_dict(bad)
    finally:
        eng.close()


def test_launch_count(tiny_net):
    net, cfg, _ = tiny_net
    eng = net.engine
    eng.reset_launch_count()
    eng.score(torch.zeros(3, 3, 224, 224, device="cuda"))
    torch.cuda.synchronize()
    # 3 embedding launches + 6 per layer + (L - 1) next-layer LayerNorms + 4 tail launches
    assert eng.launch_count == 5 * cfg.layers + 7


@pytest.mark.parametrize("cfg_name", ["tiny", "small"])
def test_cls_shortcut_is_equivalent(engine_factory, cfg_name):
    """The last-layer CLS-only shortcut (HF:685 consumes only row 0) must not change results beyond
    the rounding of one fp16 attention output."""
    from mcm_b200 import synth
    eng, sd, cfg = engine_factory(cfg_name, 5, 32)
    eng.set_text_bank(synth.synth_unit_bank(12, cfg.proj, 9))
    x = torch.from_numpy(synth.synth_images(11, 8)).cuda()
    try:
        eng.set_cls_shortcut(True)
        a_s, a_f = eng.score(x).clone(), eng.image_features(x).clone()
        eng.set_cls_shortcut(False)
        b_s, b_f = eng.score(x).clone(), eng.image_features(x).clone()
    finally:
        eng.set_cls_shortcut(True)
    torch.cuda.synchronize()
    rel = ((a_f - b_f).norm(dim=1) / b_f.norm(dim=1)).max().item()
    assert rel <= 2e-3, rel
    assert (a_s - b_s).abs().max().item() <= 2e-5


def test_uint8_ingest_is_bit_identical_to_fp32_path(tiny_net):
    """uint8 HWC pixels through the fused ToTensor + Normalize patch gather == the fp32 tensor the reference
    preprocess (utils/train_eval_util.py:27-34) would have produced, through the fp32 entry points."""
    from mcm_b200 import synth
    from oracle import clip_mcm_oracle as O
    net, cfg, bank = tiny_net
    eng = net.engine
    u8 = synth.synth_images_u8(45, 21)
    f32 = O.preprocess_u8(u8)
    ref_scores = torch.cat([eng.score(f32[s:s + 32].cuda()) for s in range(0, 45, 32)]).cpu().numpy()
    ref_feats = eng.image_features(f32[:32].cuda()).cpu()
    d_u8 = torch.from_numpy(u8).cuda()
    got = torch.cat([eng.score_u8(d_u8[s:s + 32]) for s in range(0, 45, 32)]).cpu().numpy()
    np.testing.assert_array_equal(got, ref_scores)
    assert torch.equal(eng.image_features_u8(d_u8[:32]).cpu(), ref_feats)
    for batch in (32, 7):
        np.testing.assert_array_equal(eng.score_stream_host_u8(torch.from_numpy(u8).pin_memory(), batch=batch), ref_scores)
    np.testing.assert_array_equal(eng.score_stream_host_u8(u8, batch=16), ref_scores)
    # against the oracle end to end
    ora = O.ood_scores(f32, eng_sd(tiny_net), cfg, bank, T=1, score="MCM", batch=45)
    np.testing.assert_allclose(got, ora, rtol=0, atol=1e-3)
    # other normalisation constants
    eng.set_normalization((0.5, 0.5, 0.5), (0.25, 0.5, 1.0))
    try:
        f2 = O.preprocess_u8(u8[:8], (0.5, 0.5, 0.5), (0.25, 0.5, 1.0))
        np.testing.assert_array_equal(eng.score_u8(d_u8[:8]).cpu().numpy(), eng.score(f2.cuda()).cpu().numpy())
    finally:
        eng.set_normalization(O.CLIP_MEAN, O.CLIP_STD)
    with pytest.raises(ValueError):
        eng.score_u8(d_u8[:4, :100])
    with pytest.raises(ValueError):
        eng.score_u8(d_u8[:4].float())
    with pytest.raises(ValueError):
        eng.set_normalization((0, 0, 0), (1, 0, 1))
    # a rejected call leaves the constants untouched
    np.testing.assert_array_equal(eng.score_u8(d_u8[:8]).cpu().numpy(), ref_scores[:8])


def eng_sd(tiny_net_fixture):
    from mcm_b200 import synth
    return synth.synth_vision_state_dict(tiny_net_fixture[1], 5)


def test_resize_crop_on_device_matches_oracle(tiny_net):
    """mcm_resize_crop_u8: Resize(224) + CenterCrop(224) of a ragged batch of decoded images on the device, bit-identical
    to the oracle restatement of torchvision + Pillow (pinned to them in tests/test_oracle_golden.py), and the whole
    preprocess + scoring chain equal to scoring the tensor the reference's DataLoader would have produced."""
    from oracle import clip_mcm_oracle as O
    from oracle import pil_resize_oracle as R
    from oracle.make_golden_resize import image
    net, cfg, bank = tiny_net
    eng = net.engine
    sizes = [(375, 500), (500, 375), (224, 224), (224, 300), (301, 224), (100, 160), (333, 1000), (1500, 431), (64, 64),
             (227, 229), (375, 500), (2000, 3008), (37, 1000), (500, 375)]
    imgs = [image(h, w, i) for i, (h, w) in enumerate(sizes)]
    want = np.stack([R.resize_center_crop_u8(im) for im in imgs])
    got = eng.resize_crop_u8(imgs)
    assert got.shape == (len(imgs), 224, 224, 3) and got.dtype == torch.uint8
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # twice more: the plan staging is double-buffered and reused
    np.testing.assert_array_equal(eng.resize_crop_u8(imgs[:3]).cpu().numpy(), want[:3])
    np.testing.assert_array_equal(eng.resize_crop_u8(imgs[3:9]).cpu().numpy(), want[3:9])
    scores = eng.score_images(imgs).cpu().numpy()
    ref = eng.score(O.preprocess_u8(want).cuda()).cpu().numpy()
    np.testing.assert_array_equal(scores, ref)
    assert eng.resize_crop_u8([]).shape == (0, 224, 224, 3)
    # the same through the one-call host stream (pipelined H2D + preprocess + scoring), ragged batches
    sizes_b = np.array([im.size for im in imgs], dtype=np.int64)
    offs = np.zeros(len(imgs), dtype=np.int64)
    offs[1:] = np.cumsum(sizes_b[:-1])
    packed = torch.from_numpy(np.concatenate([im.reshape(-1) for im in imgs])).pin_memory()
    hs = [im.shape[0] for im in imgs]
    ws = [im.shape[1] for im in imgs]
    for batch in (32, 5):
        np.testing.assert_array_equal(eng.score_stream_host_images(packed, offs, hs, ws, batch=batch), ref)
    with pytest.raises(ValueError):
        eng.score_stream_host_images(packed, offs + 10 ** 9, hs, ws)
    with pytest.raises(ValueError):
        eng.resize_crop_u8([np.zeros((10, 10), np.uint8)])
