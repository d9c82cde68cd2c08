"""The bench lines of record under profiles/ carry every key the measurement contract names (they are copies of what
`bench.py` printed on the GPU box; this guards the schema, not the numbers)."""
import json

import pytest

from conftest import ROOT

LINES = sorted((ROOT / "profiles").glob("r02_bench_*.json")) + [ROOT / "profiles" / "r02_reference_n1.json"]
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "e2e"]


@pytest.mark.parametrize("path", LINES, ids=lambda p: p.stem)
def test_line_schema(path):
    d = json.loads(path.read_text())
    for k in BASE:
        assert k in d, (path.name, k)
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    cb = d.get("cpu_baseline")
    if d["n_gpus"] == 1:  # the CPU baseline is timed on rank 0 at N = 1 only
        assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("port", "reference")
    if d.get("impl") == "reference":
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] == cb["value"]
        return
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
    if d["n_gpus"] > 1:
        assert d["scaling"] == "strong" and len(d["per_rank"]) == d["n_gpus"]
