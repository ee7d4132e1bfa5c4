"""The JSON line `bench.py` prints is a contract with the driver: the committed line of the last GPU run
(profiles/r2b_bench_c3_default.json, produced by `python bench.py` on a B200) must carry every key the
contract names, with consistent values.  Runs without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_default_line_has_the_contract_keys():
    d = _line("r2b_bench_c3_default.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    # value = sample-iterations per second over the timed steps
    n, it = d["config"]["samples_per_gpu"] * d["n_gpus"], d["config"]["iters_per_step"]
    assert abs(d["value"] - n * it / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference")
    k = d["clocks"]
    assert "sm_mhz" in k and "sm_max_mhz" in k and isinstance(k["reasons"], list)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"]))
    for cfg in ("c1", "c2", "c4", "c5", "c5_step"):
        assert cfg in d["configs"], cfg
    assert d["configs"]["c4"]["roofline"]["bound"] == "tensor"


def test_reference_arm_line():
    d = _line("r2b_bench_c3_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "sample-iterations/s" and d["higher_is_better"] is True
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["metric"] == _line("r2b_bench_c3_default.json")["metric"]


def test_bench_cli_parses():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout + out.stderr      # bench.py keeps stdout for its one JSON line
