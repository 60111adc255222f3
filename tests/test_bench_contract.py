"""bench.py prints exactly one JSON line with the keys the driver reads (metric, value, unit, n_gpus, steps, warmup, ms_per_step,
higher_is_better, scaling, vs_baseline, dtype, data, config.workload, clocks, e2e, gpu_launches, roofline, cpu_baseline)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
        "gpu_launches"]


def _run(args, timeout):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout                       # ONE line on stdout; everything else goes to stderr
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-seconds", "1"], timeout=600)
    for k in BASE + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "iterations" in cb["sample"]
    assert 0 < d["value"] < 10


@pytest.mark.gpu
def test_our_arm_line_on_gpu():
    d = _run(["--images", "3", "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--l2-images", "1"], timeout=900)
    for k in BASE + ["clocks", "roofline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 3 and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["value"] > 1 and d["e2e"]["value"] > 1 and d["e2e"]["value"] <= d["value"] * 1.5   # one tiny step each: timing noise
    assert d["e2e"]["h2d_bytes_per_step"] == 3 * 100 * 128 * 128 * 4 and d["e2e"]["d2h_bytes_per_step"] == 3 * 512 * 512 * 4
    assert d["gpu_launches"] >= 600                          # 300 iterations x two solve kernels, at least
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "pair_frac", "fp32_pipe_frac", "l2"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["l2"]["peak"] > 1000 and r["l2"]["resident_run"]["images_in_flight"] == 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
