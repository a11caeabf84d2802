"""Golden vectors for the Resize(224) + CenterCrop(224) step (TEST INFRASTRUCTURE): runs torchvision + Pillow
themselves (the reference's val_preprocess, utils/train_eval_util.py:29-31) on seeded random uint8 images in the
authoring container and stores the SHA-1 of every 224 x 224 x 3 result in tests/golden/resize_crop_pil.npz, so that
the oracle restatement (oracle/pil_resize_oracle.py) stays pinned where Pillow / torchvision are not importable.

    python -m oracle.make_golden_resize
"""
import hashlib
import os

import numpy as np

SIZES = [(375, 500), (500, 375), (224, 224), (224, 300), (301, 224), (225, 224), (100, 160), (160, 100), (333, 1000),
         (1500, 431), (64, 64), (227, 229), (900, 1200), (2000, 3008), (37, 1000)]
SEED = 20261017


def image(h, w, i):
    """Seeded test image: uniform noise for even i, a smooth gradient + noise (photo-like) for odd i."""
    rng = np.random.default_rng([SEED, h, w, i])
    if i % 2 == 0:
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(yy * 255.0 / max(h - 1, 1)), (xx * 255.0 / max(w - 1, 1)), ((yy + xx) % 256)], axis=-1)
    return np.clip(base + rng.normal(0, 12, size=(h, w, 3)), 0, 255).astype(np.uint8)


def main():
    from PIL import Image
    import torchvision.transforms as T
    from oracle import pil_resize_oracle as R
    tf = T.Compose([T.Resize(224), T.CenterCrop(224)])
    digests = []
    for i, (h, w) in enumerate(SIZES):
        img = image(h, w, i)
        ref = np.ascontiguousarray(np.asarray(tf(Image.fromarray(img))))
        got = R.resize_center_crop_u8(img)
        assert np.array_equal(ref, got), (h, w)
        digests.append(hashlib.sha1(ref.tobytes()).hexdigest())
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_crop_pil.npz")
    import PIL
    import torchvision
    np.savez(out, sizes=np.array(SIZES, dtype=np.int32), sha1=np.array(digests), seed=SEED,
             versions=np.array([f"Pillow {PIL.__version__}", f"torchvision {torchvision.__version__}"]))
    print(f"wrote {out}: {len(SIZES)} cases")


if __name__ == "__main__":
    main()
