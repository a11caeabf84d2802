"""Prompt fixture of BASELINE configs[2] ("ImageNet-1k ID, 1000 prompts, 80 templates averaged") -- TEST INFRASTRUCTURE.

Runs ONLY in the authoring container (needs /root/reference).  Reads, at generation time,
  * the 80 OpenAI prompt templates the reference ships but never uses (``utils/imagenet_templates.py:1-82``,
    ``openai_imagenet_template``: a list of ``lambda c: f'...{c}...'``), rendered into ``str.format`` patterns, and
  * the 1000 cleaned ImageNet class names that define the K = 1000 bank (``data/ImageNet/imagenet_class_clean.npy``,
    used by ``utils/common.py:29-34``),
and stores them with their SHA-256 digests in ``tests/golden/config3_prompts.npz`` so that the GPU box (which has no
/root/reference) can build the 1000 x 80 prompt bank.  The single template of the scoring loop
(``utils/detection_util.py:228``: ``"a photo of a {c}"``, no trailing period) is stored beside them.

Usage:  python oracle/make_golden_prompts.py
"""
import hashlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MCM_REFERENCE_ROOT", "/root/reference")


def digest(strings):
    return hashlib.sha256("\n".join(strings).encode("utf-8")).hexdigest()


def main():
    spec = importlib.util.spec_from_file_location("_ref_templates", os.path.join(REF, "utils", "imagenet_templates.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    templates = [t("{}") for t in mod.openai_imagenet_template]
    assert len(templates) == 80 and all(t.count("{}") == 1 for t in templates), len(templates)
    names = [str(n) for n in np.load(os.path.join(REF, "data", "ImageNet", "imagenet_class_clean.npy"))]
    assert len(names) == 1000
    out = os.path.join(ROOT, "tests", "golden", "config3_prompts.npz")
    np.savez_compressed(out, templates=np.array(templates), class_names=np.array(names),
                        reference_template=np.array("a photo of a {}"),
                        templates_sha256=np.array(digest(templates)), class_names_sha256=np.array(digest(names)))
    print(f"wrote {out}: {len(templates)} templates ({digest(templates)[:16]}...), {len(names)} class names "
          f"({digest(names)[:16]}...)")


if __name__ == "__main__":
    sys.exit(main())
