"""CPU checks of the measurement contract and of the binding tables: the committed bench
lines carry every key the driver reads, and every C entry point has a Python binding."""
import json
import os
import re

import pytest

import amgcl_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def load(name):
    with open(os.path.join(PROF, name)) as f:
        return json.loads(f.read())


REQUIRED = {"metric": str, "value": (int, float), "unit": str, "n_gpus": int, "steps": int, "warmup": int,
            "ms_per_step": (int, float), "higher_is_better": bool, "scaling": str, "dtype": str,
            "data": str, "config": dict, "e2e": dict, "gpu_launches": int, "clocks": dict,
            "roofline": dict}


@pytest.mark.parametrize("name", ["r1_bench_n1.json", "r1_bench_n2_p2p1.json", "r1_bench_n4_p2p1.json",
                                  "r1_bench_n8_p2p1.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = load(name)
    for key, typ in REQUIRED.items():
        assert key in d and isinstance(d[key], typ), key
    assert "vs_baseline" in d and d["vs_baseline"] is None         # BASELINE.md has no published number
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["l2"].startswith("inputs_exceed_l2")
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                 # copies are inside the timed region
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["scaling"] == ("weak" if d["n_gpus"] == 1 else "strong")
    assert d["iters"] == 27                                        # the survey's count at 256^3
    if d["n_gpus"] == 1:
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] <= 1.02
        assert r["traffic"] is None or r["traffic"] <= 1.05 * r["bytes_per_launch"]
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        p = d["parity"]
        assert p["iters_gpu"] == p["iters_ref"] and p["x_rel_err_inf"] < 1e-8


def test_reference_arm_line():
    d = load("r1_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == load("r1_bench_n1.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"] == load("r1_bench_n1.json")["config"]["workload"]


def test_every_c_entry_point_has_a_python_binding():
    """include/amgcl_b200.h is the boundary; amgcl_b200/__init__.py binds all of it (the tests
    call the product through these bindings only)."""
    header = open(os.path.join(ROOT, "include", "amgcl_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200_ctx_s", "b200_profile_entry"}
    src = open(os.path.join(ROOT, "amgcl_b200", "__init__.py")).read()
    bound = set(re.findall(r"\b(b200_[a-z0-9_]+)\b", src))
    missing = sorted(declared - bound)
    assert not missing, missing
    L = ab.lib()
    for name in declared:
        assert hasattr(L, name), name
