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


# ---- round 2: the one line carries every BASELINE config, the strong-scaling block and the measured peaks ---------------
R02 = os.path.join(ROOT, "profiles", "r02_bench_n1.json")
R02_N2 = os.path.join(ROOT, "profiles", "r02_bench_n2.json")
R02_REF = os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_r02_line_carries_all_five_configs():
    d = _line(R02)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "workloads", "strong", "peaks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["hits"]["all_planted_found_and_nothing_else"] is True and d["cpu_baseline"]["hits_equal_gpu"] is True
    w = d["workloads"]
    assert set(w) == {"c1", "c3", "c4", "c5btc", "c5eth"}
    for name in ("c1", "c3", "c5btc", "c5eth"):
        x = w[name]
        assert x["value"] > 0 and x["e2e"]["value"] > 0 and x["hits"]["all_planted_found_and_nothing_else"] is True, name
        assert 0 < x["roofline"]["frac"] < 1 and x["cpu_baseline"]["value"] > 0 and x["cpu_baseline"]["hits_equal_gpu"] in (True, None), name
        assert x["value"] > 20 * x["cpu_baseline"]["value"]
    c4 = w["c4"]
    assert c4["planted"]["found"] is True and c4["unit"] == c4["cpu_baseline"]["unit"] == "M giant steps/s"      # like for like: giant steps/s on both sides
    assert c4["value"] > 20 * c4["cpu_baseline"]["value"]
    for cur in ("c5btc", "c5eth"):
        s = d["strong"]["c5"][cur]
        assert s["planted_found_and_nothing_else"] is True and s["keys"] == 1 << 37 and 0.9 < s["efficiency"] <= 1.001
    assert d["strong"]["c4"]["planted"]["planted_found"] is True
    for k in ("lop3", "imad", "imad_wide", "lop3_imad_mix", "dfma", "imad_wide_nocarry", "imad_wide_nocarry_plus_lop3"):
        assert d["peaks"][k] > 1e12, k


def test_r02_strong_scaling_on_two_gpus():
    d1, d2 = _line(R02), _line(R02_N2)
    assert d2["n_gpus"] == 2 and d2["hits"]["all_planted_found_and_nothing_else"] is True
    assert 1.9 < d2["value"] / d1["value"] < 2.1                                   # weak scaling of the headline workload
    for cur in ("c5btc", "c5eth"):
        a, b = d1["strong"]["c5"][cur], d2["strong"]["c5"][cur]
        assert b["planted_found_and_nothing_else"] is True and a["keys"] == b["keys"]
        assert b["time_s"] < 0.55 * a["time_s"] and b["efficiency"] > 0.95          # the SAME range in half the time, clock around set-up + scan + gather
    s1, s2 = d1["strong"]["c4"]["sweep"], d2["strong"]["c4"]["sweep"]
    assert s1["giant_steps"] == s2["giant_steps"] and s2["time_s"] < 0.56 * s1["time_s"]
    assert d2["strong"]["c4"]["planted"]["planted_found"] is True


def test_r02_rerun_carries_the_undiscounted_fraction():
    """bench.py reports the roofline fraction against the work the kernel still does AND against SURVEY's undiscounted per-point figure"""
    d1, d = _line(R02), _line(os.path.join(ROOT, "profiles", "r02_bench_n1_rerun.json"))
    assert abs(d["value"] - d1["value"]) / d1["value"] < 0.02                      # another box, same build
    r = d["roofline"]
    assert r["ops_per_point_survey"] == 9950 and r["ops_per_point"] == 9950 - 3 * 162 - 480
    assert abs(r["frac_survey_ops"] - r["frac"] * 9950 / r["ops_per_point"]) < 1e-9 and r["frac"] < r["frac_survey_ops"] < 1
    for name in ("c1", "c3", "c5btc", "c5eth"):
        x = d["workloads"][name]["roofline"]
        assert x["frac"] <= x["frac_survey_ops"] < 1, name


def test_r02_four_gpus():
    d1, d4 = _line(R02), _line(os.path.join(ROOT, "profiles", "r02_bench_n4.json"))
    assert d4["n_gpus"] == 4 and d4["hits"]["all_planted_found_and_nothing_else"] is True and 3.9 < d4["value"] / d1["value"] < 4.1
    for cur in ("c5btc", "c5eth"):
        assert d4["strong"]["c5"][cur]["planted_found_and_nothing_else"] is True and d4["strong"]["c5"][cur]["efficiency"] > 0.95
    assert d4["strong"]["c4"]["sweep"]["efficiency"] > 0.9 and d4["strong"]["c4"]["planted"]["planted_found"] is True


def test_r02_eight_gpus():
    """one 8 x B200 node (gpurun --gpus 8, tools/gpu_job_n8.sh): weak-scaled headline and the strong-scaling blocks"""
    d1, d8 = _line(R02), _line(os.path.join(ROOT, "profiles", "r02_bench_n8.json"))
    assert d8["n_gpus"] == 8 and d8["hits"]["all_planted_found_and_nothing_else"] is True
    assert 7.8 < d8["value"] / d1["value"] < 8.2
    for cur in ("c5btc", "c5eth"):
        a, b = d1["strong"]["c5"][cur], d8["strong"]["c5"][cur]
        assert b["planted_found_and_nothing_else"] is True and a["keys"] == b["keys"] and b["time_s"] < a["time_s"] / 7.5 and b["efficiency"] > 0.95
    s1, s8 = d1["strong"]["c4"]["sweep"], d8["strong"]["c4"]["sweep"]
    assert s1["giant_steps"] == s8["giant_steps"] and s8["time_s"] < s1["time_s"] / 7 and s8["efficiency"] > 0.9
    assert d8["strong"]["c4"]["planted"]["planted_found"] is True


def test_r02_reference_arm_agrees_with_the_inline_cpu_baseline():
    r, d = _line(R02_REF), _line(R02)
    assert r["impl"] == "reference" and r["e2e"]["h2d_bytes_per_step"] == 0
    assert abs(r["value"] - d["cpu_baseline"]["value"]) / d["cpu_baseline"]["value"] < 0.03
