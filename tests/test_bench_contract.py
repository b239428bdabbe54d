"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm (the numpy restatement of the
reference's CPU path) prints exactly one JSON line with the keys the driver reads, and the B200 arm fails loudly
(no CPU fallback) when there is no device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench("--impl", "reference", "--config", "c1", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "sequences/s" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == pytest.approx(d["value"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == pytest.approx(d["value"])


def test_b200_arm_without_a_gpu_fails_loudly():
    import ctypes
    try:
        n = ctypes.CDLL("libcuda.so.1").cuInit(0)
    except OSError:
        n = 1
    if n == 0:
        pytest.skip("a CUDA device is present")
    r = run_bench("--config", "c1", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "NOGPU" in r.stderr or "no CUDA device" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
