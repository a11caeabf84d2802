"""The C-ABI library loads and exports every symbol include/mcm_b200.h declares (no GPU needed),
and the product path fails loudly -- never falls back -- without a CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mcm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from mcm_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(build.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mcm_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.load().mcm_abi_version() == _lib.ABI_VERSION


def test_flops_entry_point_needs_no_gpu():
    from mcm_b200 import _lib, synth
    c = synth.CFGS["ViT-B/16"]
    cc = _lib.McmConfig(c.image_size, c.patch, c.width, c.layers, c.heads, c.mlp, c.proj, c.eps, 1, 0)
    assert _lib.load().mcm_flops_per_image(ctypes.byref(cc), 1000) == pytest.approx(35.128e9, rel=2e-4)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from mcm_b200 import _lib, synth
    from mcm_b200.engine import McmEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        McmEngine(synth.CFGS["tiny"])
    c = synth.CFGS["tiny"]
    cc = _lib.McmConfig(c.image_size, c.patch, c.width, c.layers, c.heads, c.mlp, c.proj, c.eps, 1, 0)
    h = ctypes.c_void_p()
    rc = _lib.load().mcm_create(ctypes.byref(cc), ctypes.byref(h))
    assert rc == _lib.ECUDA and not h.value
    assert b"no CPU path" in _lib.load().mcm_last_error(None)


def test_create_rejects_bad_shapes():
    from mcm_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    for bad in [dict(width=100), dict(heads=5), dict(mlp=1000), dict(proj=6), dict(patch=15), dict(layers=0)]:
        kw = dict(image_size=224, patch=16, width=768, layers=12, heads=12, mlp=3072, proj=512, eps=1e-5,
                  max_batch=8, device=0)
        kw.update(bad)
        cc = _lib.McmConfig(*[kw[f[0]] for f in _lib.McmConfig._fields_])
        assert lib.mcm_create(ctypes.byref(cc), ctypes.byref(h)) in (_lib.EINVAL, _lib.EUNSUPPORTED), bad
        assert not h.value
