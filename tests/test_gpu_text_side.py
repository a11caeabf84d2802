"""The text side of the drop-in, end to end on the device:

  * seam 2 construction: ``set_model_clip`` / ``wrap_clip_model`` (``utils/train_eval_util.py:15-36``) build the B200 ``net``
    from a HuggingFace ``CLIPModel``; ``get_ood_scores_clip`` is then called WITHOUT a pre-encoded bank, so the
    tokenizer -> text-tower branch that mirrors ``utils/detection_util.py:216,228-231`` runs, and the scores are compared
    with the goldens the UNMODIFIED reference produced on the same (random-init) model and fake tokenizer;
  * BASELINE configs[2]: the K = 1000 x 80-template prompt bank (strings of ``utils/imagenet_templates.py:1-82`` and
    ``data/ImageNet/imagenet_class_clean.npy``, frozen with their digests in tests/golden/config3_prompts.npz by
    oracle/make_golden_prompts.py) built by ``mcm_b200.text_bank.build_text_bank`` and scored against on the engine.
"""
import hashlib
import os
import types

import numpy as np
import pytest
import torch

from helpers import ListLoader, golden_inputs, report

pytestmark = pytest.mark.gpu


def _digest(strings):
    return hashlib.sha256("\n".join(strings).encode("utf-8")).hexdigest()


@pytest.mark.parametrize("case,via", [("tiny_text_k10", "wrap_clip_model"), ("b16_text_k10_cfg1", "set_model_clip")])
def test_scores_through_tokenizer_and_text_tower(case, via, golden_dir, monkeypatch):
    import transformers
    from mcm_b200 import detection_util as DU
    from mcm_b200 import train_eval_util as TU
    from oracle import reference_shims as R
    from oracle.make_golden import IMAGENET10
    z = np.load(os.path.join(golden_dir, case + ".npz"))
    cfg, sd, _protos, id_imgs, ood_imgs = golden_inputs(z)
    # the same random-init HF CLIPModel the reference ran on when the fixture was made (vision weights = sd, text tower
    # seeded); the 4.x/5.x API shim of the model is irrelevant here: B200ClipNet unwraps either return type itself
    model = R.make_shim_clip(cfg, sd, seed=int(z["wseed"]))
    type(model).bank_override = None
    labels = IMAGENET10[: int(z["K"])]
    monkeypatch.setattr(DU, "CLIPTokenizer", R.FakeTokenizer)          # no tokenizer vocabulary offline
    args = types.SimpleNamespace(CLIP_ckpt=str(z["cfg"]), model="CLIP", score="MCM", T=int(z["T"]), batch_size=128, gpu=0)
    if via == "set_model_clip":
        monkeypatch.setattr(transformers.CLIPModel, "from_pretrained", classmethod(lambda cls, name, *a, **k: model))
        net, preprocess = TU.set_model_clip(args)
        assert args.ckpt == "openai/clip-vit-base-patch16"             # utils/train_eval_util.py:19-22
        names = [type(t).__name__ for t in preprocess.transforms]
        assert names == ["Resize", "CenterCrop", "ToTensor", "Normalize"]   # :29-34
        assert tuple(preprocess.transforms[3].mean) == TU.CLIP_MEAN and tuple(preprocess.transforms[3].std) == TU.CLIP_STD
    else:
        net = TU.wrap_clip_model(args, model)
        assert args.ckpt == str(z["cfg"])
    net = net.eval()
    eng = net.engine
    try:
        assert net.text_bank is None and eng.K == 0                    # nothing pre-encoded
        for sc in [str(s) for s in z["scores"]]:
            args.score = sc
            key = sc.replace("-", "_")
            got_in = DU.get_ood_scores_clip(args, net, ListLoader(id_imgs, 96), labels, in_dist=True)
            got_out = DU.get_ood_scores_clip(args, net, ListLoader(ood_imgs, 50), labels)
            ref_in, ref_out = z[f"ref_in_{key}"], z[f"ref_out_{key}"]
            err = max(np.abs(got_in - ref_in).max(), np.abs(got_out - ref_out).max())
            report("text_side", dict(case=case, via=via, score=sc, max_abs_err=float(err)))
            assert got_in.dtype == np.float32 and got_in.shape == ref_in.shape and got_out.shape == ref_out.shape
            assert err <= 1e-3, (case, sc, err)
        assert eng.K == int(z["K"])
        # the bank the engine ended up with is the reference's: text tower on "a photo of a {c}", rows normalised (:228-231)
        tok = R.FakeTokenizer()(DU.prompt_texts(labels))
        with torch.no_grad():
            tf = model.get_text_features(input_ids=tok["input_ids"], attention_mask=tok["attention_mask"])
        tf = tf / tf.norm(dim=-1, keepdim=True)
        assert np.abs(tf.numpy() - z["bank"]).max() <= 1e-6
    finally:
        eng.close()


def test_config3_bank_of_80_templates_end_to_end(golden_dir):
    """ViT-B/16, K = 1000 ImageNet classes x the 80 OpenAI templates, averaged: builder vs an independent fp64 restatement
    of the ensemble, then engine vs oracle scores / AUROC / FPR95 on ID and OOD streams with that bank."""
    from mcm_b200 import detection_util as DU
    from mcm_b200 import synth
    from mcm_b200.engine import B200ClipNet, McmEngine
    from mcm_b200.text_bank import build_text_bank, render
    from oracle import clip_mcm_oracle as O
    from oracle import reference_shims as R
    z = np.load(os.path.join(golden_dir, "config3_prompts.npz"))
    templates = [str(t) for t in z["templates"]]
    names = [str(n) for n in z["class_names"]]
    assert len(templates) == 80 and len(names) == 1000
    assert _digest(templates) == str(z["templates_sha256"]) and _digest(names) == str(z["class_names_sha256"])
    assert templates[0] == "a bad photo of a {}." and names[0] == "tench"     # utils/imagenet_templates.py:2

    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = synth.CFGS["ViT-B/16"]
    sd = synth.synth_vision_state_dict(cfg, 5)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    model = R.make_shim_clip(cfg, sd, seed=5, text_layers=2, text_width=128, text_heads=2).cuda()
    type(model).bank_override = None
    eng = McmEngine.from_state_dict(sd, cfg, max_batch=256)
    try:
        net = B200ClipNet(eng, text_model=model).eval()
        tokenizer = R.FakeTokenizer()
        bank = build_text_bank(net, tokenizer, names, templates=templates, chunk=4000, cache_key="cfg3-test")
        assert bank.shape == (1000, cfg.proj) and bank.dtype == torch.float32
        assert torch.allclose(bank.norm(dim=-1), torch.ones(1000), atol=1e-5)
        # independent restatement on a slice of classes, fp64: per-template normalise -> mean -> re-normalise
        sl = list(range(0, 1000, 97))
        acc = torch.zeros((len(sl), cfg.proj), dtype=torch.float64)
        with torch.no_grad():
            for t in templates:
                tok = tokenizer([render(t, names[i]) for i in sl])
                f = model.get_text_features(input_ids=tok["input_ids"].cuda(), attention_mask=tok["attention_mask"].cuda()).double().cpu()
                acc += f / f.norm(dim=-1, keepdim=True)
        ens = acc / len(templates)
        ens = ens / ens.norm(dim=-1, keepdim=True)
        assert (bank[sl].double() - ens).abs().max().item() <= 2e-6
        # a second call is served from the cache (one encoding per label set)
        again = build_text_bank(net, tokenizer, names, templates=templates, chunk=4000, cache_key="cfg3-test")
        assert torch.equal(again, bank)

        # engine vs oracle with this bank: 2 000 ID + 2 000 OOD images
        net_b = B200ClipNet(eng, text_bank=bank).eval()
        args = types.SimpleNamespace(ckpt="cfg3", model="CLIP", score="MCM", T=1, batch_size=256)
        bank_gpu = bank.cuda()
        bank_gpu = bank_gpu / bank_gpu.norm(dim=-1, keepdim=True)
        ident = synth.synth_images(2000, 11)
        ood = synth.synth_images(2000, 12, mean=0.3, std=1.5)
        with torch.no_grad():
            ref_in = O.ood_scores(torch.from_numpy(ident).cuda(), sd_gpu, cfg, bank_gpu, T=1, score="MCM", batch=125)
            ref_out = O.ood_scores(torch.from_numpy(ood).cuda(), sd_gpu, cfg, bank_gpu, T=1, score="MCM", batch=125)
        m_ref = O.get_measures(-ref_in, -ref_out)
        std = float(np.concatenate([ref_in, ref_out]).std())
        for precision in ("fp16", "split"):
            eng.set_precision(precision)
            got_in = DU.get_ood_scores_clip(args, net_b, ListLoader(ident, 256), names, in_dist=True)
            got_out = DU.get_ood_scores_clip(args, net_b, ListLoader(ood, 256), names)
            m_got = DU.get_measures(-got_in, -got_out)
            err = float(max(np.abs(got_in - ref_in).max(), np.abs(got_out - ref_out).max()))
            report("config3", dict(precision=precision, score_std=std, max_abs_err=err, auroc_ref=float(m_ref[0]), auroc=float(m_got[0]),
                                   fpr_ref=float(m_ref[2]), fpr=float(m_got[2])))
            assert err <= 1e-3
            if precision == "split":
                # fp32-class arithmetic: the bars proper.  (A random-init text tower gives a bank unrelated to the images, so
                # the scores differ between images by ~1e-6 only and fp16 operand rounding re-orders them: for the fp16 mode
                # this stream is a noise-ordering test, the designed-margin streams are in test_gpu_parity_k1000.py.)
                assert err <= 0.02 * std, (err, std)
                assert abs(m_got[0] - m_ref[0]) <= 5e-4, (m_got, m_ref)
                assert abs(m_got[2] - m_ref[2]) <= 5e-4 + 1.0 / len(ref_out), (m_got, m_ref)
            else:
                assert abs(m_got[0] - m_ref[0]) <= 0.05, (m_got, m_ref)
    finally:
        eng.set_precision("fp16")
        eng.close()
