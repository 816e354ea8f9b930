"""The bench.py JSON lines committed under profiles/ carry every key of the measurement contract (own arm and reference arm).
CPU-only: it reads the recorded lines, it does not run the bench."""
import json
import os

import pytest

from conftest import ROOT


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r01_bench_n1.json", "r02_bench_n1.json"])
def test_own_arm_line_has_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["value"] > 0 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("rnd", ["r01", "r02"])
def test_reference_arm_line_has_the_contract_keys(rnd):
    d = _line(f"{rnd}_bench_reference_n1.json")
    own = _line(f"{rnd}_bench_n1.json")
    assert d["impl"] == "reference" and d["metric"] == own["metric"] and d["unit"] == own["unit"]
    assert d["higher_is_better"] == own["higher_is_better"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["value"] == d["value"] and c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
