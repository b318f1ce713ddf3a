"""bench.py's output contract (the parts that can run without a GPU): exactly one JSON line on stdout with the
keys the driver reads; the reference arm times the CPU port and says so; the GPU arm fails loudly without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-proteins", "3000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("kmer occurrences/s") and d["unit"] == "occurrences/s" and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["data"] == "synthetic" and d["dtype"] == "u64"
    assert d["config"]["workload"] == "config2" and "sample" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "proteins" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU arm is exercised by the round-end bench")
    r = run_bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--workload", "config1")
    assert r.returncode != 0
    assert "sigk_create" in (r.stderr + r.stdout) or "CUDA" in (r.stderr + r.stdout) or "cuda" in (r.stderr + r.stdout)
