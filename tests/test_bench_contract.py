"""bench.py's command-line contract, as far as it can be checked without a GPU: the reference arm's JSON line
(CPU oracle port of utils/detection_util.py:209-249), what ranks > 0 do under torchrun, and that the GPU arm refuses to
run -- loudly, no CPU fallback -- when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    # ViT-B/32 keeps the CPU work of the two 16-image steps to a few seconds; the line's layout does not depend on the model
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--model", "ViT-B/32", "--K", "100"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "images/s"
    assert d["metric"] == "images/sec MCM-scored (ViT-B/32, K=100)"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "16 images" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None      # BASELINE.md holds no published number for this metric


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"}, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_gpu_arm_fails_loudly_without_cuda():
    r = _run(["--steps", "1", "--warmup", "0"], timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
