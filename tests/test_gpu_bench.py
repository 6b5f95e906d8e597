"""The benchmark's own contract, on the GPU: `python bench.py` prints exactly one JSON line with the
keys the driver reads (the short form: no extras, no CPU legs)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_prints_one_json_line(drrt):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                          "--no-extra", "--no-cpu"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["gpu_launches"] == 2                      # one range_cta_kernel launch per timed step
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["traffic"] is None or r["traffic_source"].startswith("profiles/")
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 24 * 65536 and e["d2h_bytes_per_step"] > 2e9
    assert e["value"] < d["value"]                     # the host call includes the copies
    assert e["ids_only"]["value"] > e["rows_f32"]["value"] > e["value"]
    assert 3000 < d["config"]["hits_per_query"] < 4000
