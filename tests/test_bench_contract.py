"""The reference arm of bench.py runs entirely on the CPU, so its side of the driver's contract is checked here:
one JSON line with the keys the driver reads, `impl: reference`, and no work on ranks other than 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "ERP pairs/s" and d["unit"] == "pairs/s" and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    env = dict(os.environ, SBA_BENCH_PAIRS_PER_STEP="8")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--cpu-budget", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "e2e_reference_order", "gpu_launches", "roofline", "cpu_baseline", "clocks", "per_rank_ms", "run"):
        assert key in d, key
    assert "impl" not in d and d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3 and d["scaling"] == "weak"     # --warmup honoured exactly
    pps = d["run"]["pairs_per_step_per_gpu"]
    assert pps == 8 and d["value"] > 100 and abs(d["ms_per_step"] * d["value"] - 1000.0 * pps) < 1.0      # pairs/s and ms per step agree
    assert len(d["per_rank_ms"]["device_resident"]) == 1 and abs(d["per_rank_ms"]["device_resident"][0] - d["ms_per_step"] * 6) < 1e-6
    e, ero = d["e2e"], d["e2e_reference_order"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 50_000_000 * pps and e["d2h_bytes_per_step"] > 0
    assert 0 < ero["value"] <= e["value"] * 1.05 and ero["d2h_bytes_per_step"] > 33_000_000 * pps          # both strips come back
    assert d["gpu_launches"] >= 6 * pps * 6
    rf = d["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["frac"] < 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and isinstance(cb["sample"], str)
    ck = d["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(ck["reasons"]))
    assert "workload" in d["config"] and "model" not in d["config"]


def test_both_arms_share_metric_and_config():
    """The driver compares the two arms' `config` dicts and metric strings: they come from the same constants."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": CONFIG') == 2 and src.count('"metric": METRIC') == 2
    assert b.CONFIG["workload"].startswith("C2") and b.PAIRS_PER_STEP >= 1
