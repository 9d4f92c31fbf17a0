"""bench.py contract on a box without a GPU: the reference arm prints exactly one JSON line with the
required keys, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, **env):
    e = dict(os.environ, **env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    out = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-n", "12"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "DOF/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert abs(d["ms_per_step"] - 1e3 * d["cpu_baseline"]["seconds_per_solve"]) < 1e-6


def test_reference_arm_other_ranks_do_nothing():
    out = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample-n", "12"],
              RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = run(["-n", "8", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
