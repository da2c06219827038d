"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys,
the product arm refuses to run without a CUDA device (no CPU fallback), the clock sampler's window logic."""
import json
import os
import subprocess
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-frames", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "frames/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["config"]["workload"] == "cityscapes_256x512_c64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["vs_baseline"] is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_product_arm_has_no_cpu_fallback():
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_keeps_rows_of_the_timed_region():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler([0, 1])
    t = time.perf_counter()
    row = lambda mhz, cap="Not Active": [str(mhz), "1965", "500.0", "Not Active", "Not Active", "Not Active", cap]  # noqa: E731
    s.rows = [(t - 5.0, row(300)), (t + 0.01, row(1950)), (t + 0.02, row(1965, "Active")), (t + 9.0, row(210))]
    s.t0, s.t1 = t, t + 0.03
    out = s.summary()
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["gpus"] == 2 and "window" not in out
    # a region shorter than one sampling period falls back to the nearest rows and says so
    s.rows = [(t - 0.2, row(1965)), (t + 0.25, row(1965))]
    out = s.summary()
    assert out["samples"] == 2 and out["window"] == "+-0.3 s"
    # no sampler on the other ranks
    assert bench.ClockSampler([]).summary() is None
