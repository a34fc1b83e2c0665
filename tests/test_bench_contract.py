"""CPU: the bench.py contract the driver relies on — the reference arm falls back to the CPU oracle when no GPU / harness is
usable and still prints ONE well-formed JSON line; our arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def run_bench(*args, env_extra=None):
    env = dict(os.environ, XM_BENCH_CAMERAS="60", **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_on_cpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the reference arm runs the compiled reference harness instead")
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "xm_tcg_iterations_per_sec" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_reference_arm_is_silent_on_non_zero_ranks():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env_extra={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = run_bench("--steps", "1", "--warmup", "0")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
