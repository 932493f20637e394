"""CPU: the bench lines committed under profiles/ carry every key of the driver's contract (bench.py docstring), with
values that are consistent with each other.  (The lines themselves are produced on the B200 by bench.py / bench_bsgs.py.)"""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_*.json")))


def test_there_are_bench_lines():
    names = {os.path.basename(p) for p in LINES}
    assert {"r01_bench_c2_n1.json", "r01_bench_c1_n1.json", "r01_bench_c3_n1.json", "r01_bench_c5btc_n1.json", "r01_bench_c5eth_n1.json",
            "r01_bench_c4_bsgs_k512.json", "r01_bench_c2_n2.json", "r01_bench_c2_n8.json"} <= names


@pytest.mark.parametrize("path", LINES, ids=os.path.basename)
def test_line_follows_the_contract(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["value"] > 0 and d["gpu_launches"] > 0
    assert isinstance(d["config"].get("workload"), str) and "model" not in d["config"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["value"] > 0 and c["cores"] >= 1 and c["kind"] in ("reference", "port") and c["sample"]
        assert d["value"] > 20 * c["value"] or d["unit"] != c["unit"]       # the GPU path is not a disguised CPU fallback
    # whole-job value = units / time
    if "keys_per_step_per_gpu" in d["config"]:
        per_step = d["config"]["keys_per_step_per_gpu"] * d["n_gpus"]
        assert abs(d["value"] - per_step / (d["ms_per_step"] * 1e-3) / 1e6) / d["value"] < 1e-3
