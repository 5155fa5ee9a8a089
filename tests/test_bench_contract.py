"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line carrying the keys
the driver reads, and the GPU arm refuses to run (loudly, non-zero) without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from common import ROOT


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"GQ_REF_SAMPLE": "2000"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "quasimap reads/sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("config2") and d["config"]["kmer_size"] == 10
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""
