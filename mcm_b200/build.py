"""In-tree build of the sm_100a extension (``mcm_b200/_C/libmcm_b200.so``).

nvcc cross-compiles for sm_100a without a GPU, so this runs in the authoring
container; the built ``.so`` is git-ignored but travels to the GPU box with the
repo snapshot.  Nothing here falls back to another backend: without nvcc the
build fails, without the ``.so`` the package refuses to run.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OUT_DIR = os.path.join(ROOT, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libmcm_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-ldl",
]


def _sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(ROOT), "include", "mcm_b200.h"))
    return deps


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_variant(name: str, defines) -> str:
    """A/B builds for GPU experiments: ``_C/libmcm_b200_<name>.so`` compiled with extra ``-D`` flags;
    tools select it with ``mcm_b200._lib.use_library(path)`` before the first engine is created."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, f"libmcm_b200_{name}.so")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out, os.path.join(CSRC, "engine.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/engine.cu`` (which includes every kernel header) into the shared library."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: the mcm_b200 CUDA extension cannot be built (there is no CPU path)")
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "engine.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
