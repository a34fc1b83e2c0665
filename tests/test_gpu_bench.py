"""GPU: bench.py's contract on a reduced problem (4000 cameras, 1.15 GB of Q — the large-problem code path: device generator,
time-capped steps, pinned slab upload) — both arms print ONE JSON line with the keys the driver reads, the same `config`, and
consistent numbers."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
ENV = dict(XM_BENCH_CAMERAS="4000", XM_BENCH_STEP_SECONDS="0.4", XM_BENCH_EXTRAS="0", XM_BENCH_FULL_SOLVE="0")


def run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=dict(os.environ, **ENV), timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_both_arms_print_the_contract_line():
    ours = run("--steps", "2", "--warmup", "1", "--cpu-seconds", "1")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in ours, key
    assert ours["metric"] == "xm_tcg_iterations_per_sec" and ours["dtype"] == "f64" and ours["n_gpus"] == 1 and ours["gpu_launches"] == 2
    assert ours["value"] > 0 and ours["e2e"]["value"] > 0 and ours["e2e"]["value"] <= 1.05 * ours["value"]
    assert ours["e2e"]["h2d_bytes_per_step"] >= 8 * (3 * 4000) ** 2 and ours["e2e"]["d2h_bytes_per_step"] > 0
    rf = ours["roofline"]
    assert rf["bound"] == "hbm" and 0.3 < rf["frac"] < 1.2 and abs(rf["achieved"] / rf["peak"] - rf["frac"]) < 1e-9
    assert ours["cpu_baseline"]["kind"] == "port" and ours["cpu_baseline"]["cores"] >= 1
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "xm_ref_harness")):
        pytest.skip("reference harness not built")
    ref = run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"] == ours["config"]                      # the driver's same_config check
    assert ref["value"] > 0 and ref["e2e"]["value"] <= ref["value"] and ref["cpu_baseline"]["kind"] == "reference"
    assert ours["value"] > ref["value"]                          # the point of the exercise
