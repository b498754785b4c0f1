"""CPU: the committed bench lines (profiles/r02_final_*) carry every key the measurement contract names, with
consistent arithmetic — a guard against bench.py drifting away from the contract (the lines themselves are what a
B200 printed; nothing here runs a kernel)."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parent.parent / "profiles"


def _line(name):
    p = PROFILES / name
    if not p.exists():
        pytest.skip(f"{name} not committed")
    return json.loads(p.read_text().strip().splitlines()[-1])


def test_gpu_arm_line_has_the_contract_keys():
    d = _line("r02_final_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "l2" in d["config"]
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-2 * d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3 and 0 < r["frac"] <= 1
    assert abs(r["achieved"] - r["algorithmic_flops_per_launch"] / (r["us_per_launch"] * 1e-6) / 1e12) < 0.01 * r["achieved"]
    h = d["roofline_hbm"]
    assert h["bound"] == "hbm" and 0 < h["frac"] <= 1 and h["traffic"] >= 0.99 * 8192 * 30000 * 4
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    assert d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"] > 0


def test_reference_arm_line_has_the_contract_keys():
    g, d = _line("r02_final_bench_1gpu.json"), _line("r02_final_bench_reference_arm.json")
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == g[k], k
    assert d["config"]["workload"] == g["config"]["workload"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("name,world", [("r02_final_bench_2gpu.json", 2), ("r02_final_bench_8gpu.json", 8)])
def test_multi_gpu_lines_report_identical_replicas(name, world):
    d = _line(name)
    assert d["n_gpus"] == world and d["config"]["global_batch"] == 8192 * world
    assert d["dp_params_identical_across_ranks"] is True
    assert d["dp_overlap"]["single_graph"] is True and d["dp_overlap"]["max_rel_diff_vs_plain_allreduce"] < 1e-5
    assert abs(d["value"] - world * 1e3 / d["ms_per_step"]) < 1e-2 * d["value"]
