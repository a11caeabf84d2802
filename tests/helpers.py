"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT_DIR = os.path.join(ROOT, "gpurun_out")


class ListLoader:
    """Minimal DataLoader stand-in: iterable of (images, labels) batches with ``.dataset``."""

    def __init__(self, images, batch_size, pin=False):
        self.images = torch.as_tensor(images)
        if pin and torch.cuda.is_available():
            self.images = self.images.pin_memory()
        self.batch_size = int(batch_size)
        self.dataset = range(self.images.shape[0])

    def __len__(self):
        return -(-self.images.shape[0] // self.batch_size)

    def __iter__(self):
        for s in range(0, self.images.shape[0], self.batch_size):
            x = self.images[s:s + self.batch_size]
            yield x, torch.zeros(x.shape[0], dtype=torch.long)


def make_args(T=1, score="MCM", ckpt="synthetic", batch_size=64):
    return types.SimpleNamespace(T=T, score=score, ckpt=ckpt, model="CLIP", batch_size=batch_size)


def report(name, payload):
    """Append measured errors to gpurun_out/parity_report.jsonl (brought back from the GPU box)."""
    try:
        os.makedirs(REPORT_DIR, exist_ok=True)
        with open(os.path.join(REPORT_DIR, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"name": name, **payload}) + "\n")
    except OSError:
        pass


def golden_inputs(z):
    """Regenerate the seeded inputs of a golden fixture (same code path that made it)."""
    from oracle.make_golden import build_inputs
    spec = dict(cfg=str(z["cfg"]), kind=str(z["kind"]), K=int(z["K"]), n_id=int(z["n_id"]), n_ood=int(z["n_ood"]),
                wseed=int(z["wseed"]), noise=float(z["noise"]), T=int(z["T"]))
    return build_inputs(spec)
