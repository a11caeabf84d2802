"""CPU restatement of the FIRST two steps of the reference preprocess -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``val_preprocess`` (``utils/train_eval_util.py:29-34``) is ``Resize(224) -> CenterCrop(224) -> ToTensor -> Normalize``
applied to PIL images by the DataLoader workers.  The arithmetic of the first two lives in third-party
dependencies the reference neither vendors nor pins (``README.md``: ``torchvision``; Pillow underneath):

* ``torchvision.transforms.functional.resize`` with an int size (installed: torchvision 0.26): shorter edge ->
  ``size``, longer edge -> ``int(size * long / short)``; bilinear with antialias == ``PIL.Image.resize(..., BILINEAR)``;
* Pillow's ``ImagingResample`` (``src/libImaging/Resample.c``, installed: Pillow 12.2): separable two-pass
  resampling, horizontal pass first; per output pixel a triangle filter of support ``max(scale, 1)`` whose double
  coefficients are normalised, turned into 22-bit fixed point (``(int)(0.5 + k * 2**22)``), accumulated in int32
  from ``1 << 21`` and shifted back / clipped to 8 bits -- after EACH pass;
* ``torchvision.transforms.functional.center_crop``: ``top = int(round((h - 224) / 2.0))`` (Python's round-half-even),
  likewise ``left``.

Pinned bit-for-bit against torchvision + Pillow themselves on random images of many sizes (including up-scaling and
strong down-scaling) by ``tests/test_oracle_golden.py::test_resize_oracle_matches_torchvision``; committed golden
vectors in ``tests/golden/resize_crop_*.npz`` (made by ``oracle/make_golden_resize.py``) carry the pin to the GPU box.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2      # Resample.c


def resized_size(h: int, w: int, size: int = 224):
    """torchvision ``_compute_resized_output_size`` for a single int size: (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_offsets(new_h: int, new_w: int, size: int = 224):
    """torchvision ``center_crop``: (top, left), Python round() = round half to even."""
    return int(round((new_h - size) / 2.0)), int(round((new_w - size) / 2.0))


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c ``precompute_coeffs`` (box = the whole axis, bilinear filter) followed by ``normalize_coeffs_8bpc``:
    ``bounds[out_size, 2]`` = (first source index, count) and ``kk[out_size, ksize]`` int32 fixed-point weights."""
    scale = float(in_size) / out_size           # (double)(in1 - in0) / outSize with in0 = 0, in1 = inSize (floats)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale                 # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)      # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            wgt = 1.0 - a if a < 1.0 else 0.0
            k[x] = wgt
            ww += wgt
        if ww != 0.0:
            for x in range(xmax):
                k[x] /= ww
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)      # arithmetic shift, then clip8_lookups


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One resampling pass along ``axis`` (1 = horizontal, 0 = vertical) of a uint8 [H, W, 3] image."""
    src = img.astype(np.int64)
    n = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = n
    out = np.empty(shape, dtype=np.uint8)
    for xx in range(n):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        k = kk[xx, :cnt].astype(np.int64)
        if axis == 1:
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, xmin:xmin + cnt, :], k, axes=([1], [0]))
            out[:, xx, :] = _clip8(acc)
        else:
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[xmin:xmin + cnt, :, :], k, axes=([0], [0]))
            out[xx, :, :] = _clip8(acc)
    return out


def resize_bilinear_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """``PIL.Image.resize((new_w, new_h), BILINEAR)`` of a uint8 [H, W, 3] image (``ImagingResample``: the
    horizontal pass runs first, over the source rows the vertical pass will use, and is rounded to uint8)."""
    h, w, _ = img.shape
    out = img
    if new_w != w:
        bh, kh = precompute_coeffs(w, new_w)
        out = _pass(out, bh, kh, axis=1)
    if new_h != h:
        bv, kv = precompute_coeffs(h, new_h)
        out = _pass(out, bv, kv, axis=0)
    return out.copy() if out is img else out


def resize_center_crop_u8(img: np.ndarray, size: int = 224) -> np.ndarray:
    """``CenterCrop(size)(Resize(size)(pil_image))`` as a uint8 [size, size, 3] array."""
    h, w, _ = img.shape
    new_h, new_w = resized_size(h, w, size)
    top, left = crop_offsets(new_h, new_w, size)
    r = resize_bilinear_u8(img, new_h, new_w)
    return np.ascontiguousarray(r[top:top + size, left:left + size, :])
