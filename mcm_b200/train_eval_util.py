"""B200 counterpart of the model half of the reference's ``utils/train_eval_util.py:15-36``.

``set_model_clip(args)`` there maps ``--CLIP_ckpt`` to a HuggingFace hub id, stores it in
``args.ckpt``, loads ``CLIPModel`` onto the GPU and returns ``(model, val_preprocess)``.  Here the
vision tower + projection of that same ``CLIPModel`` are packed into an :class:`McmEngine`; the
text tower stays the HF module (it runs once per label set).
"""
from __future__ import annotations

from .engine import B200ClipNet, McmEngine
from .synth import CFGS

# utils/train_eval_util.py:19-21
MODEL_CHECKPOINTS = {
    "ViT-B/32": "openai/clip-vit-base-patch32",
    "ViT-B/16": "openai/clip-vit-base-patch16",
    "ViT-L/14": "openai/clip-vit-large-patch14",
}

# utils/train_eval_util.py:27-28 -- the constants of the torchvision preprocess (data side, not timed)
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def wrap_clip_model(args, hf_model, max_batch=None, device=0) -> B200ClipNet:
    """Build the B200 ``net`` from an already-loaded HF ``CLIPModel`` (any transformers version)."""
    name = getattr(args, "CLIP_ckpt", "ViT-B/16")
    if name not in CFGS:
        raise ValueError(f"unknown --CLIP_ckpt {name!r}")
    args.ckpt = MODEL_CHECKPOINTS.get(name, name)           # side effect the reference relies on, :22
    mb = int(max_batch or min(int(getattr(args, "batch_size", 256)), 512))
    engine = McmEngine.from_state_dict(hf_model.state_dict(), CFGS[name], max_batch=mb, device=device)
    return B200ClipNet(engine, text_model=hf_model)


def set_model_clip(args, max_batch=None, device=None):
    """Reference signature: returns ``(net, val_preprocess)``; needs the checkpoint to be loadable
    by ``transformers`` (network or local cache) exactly like the reference."""
    from transformers import CLIPModel
    name = args.CLIP_ckpt
    args.ckpt = MODEL_CHECKPOINTS[name]
    model = CLIPModel.from_pretrained(args.ckpt).eval()
    net = wrap_clip_model(args, model, max_batch=max_batch, device=getattr(args, "gpu", 0) if device is None else device)
    try:
        import torchvision.transforms as transforms
        val_preprocess = transforms.Compose([
            transforms.Resize(224), transforms.CenterCrop(224), transforms.ToTensor(),
            transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])
    except Exception:  # pragma: no cover
        val_preprocess = None
    return net, val_preprocess
