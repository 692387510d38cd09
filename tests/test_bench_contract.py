"""The bench line keeps its contract: keys of the b200 arm (checked on the committed line of the final tree, which the
GPU box produced) and of the reference arm (run here on a tiny sample)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRIC = "rq_forward_encode_decode_tokens_per_sec"   # BASELINE.json: "RQ encode+decode tokens/sec per GPU ..."; both arms use this name
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config"]


def _last_line(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


def test_committed_b200_line_has_every_contract_key():
    path = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[3-9]?_bench.json")))[-1]
    d = _last_line(path)
    for k in BASE + ["roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches", "parity"]:
        assert k in d, (path, k)
    assert d["metric"] == METRIC and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert set(["bound", "achieved", "peak", "unit", "frac", "traffic"]) <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] == d["steps"] and d["parity"]["failures"] == 0
    c = d["cpu_baseline"]
    assert set(["value", "unit", "cores", "kind", "sample"]) <= set(c) and c["kind"] in ("port", "reference")
    assert set(["sm_mhz", "sm_max_mhz", "reasons"]) <= set(d["clocks"])


def test_reference_arm_prints_the_same_metric_on_a_bounded_sample():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for k in BASE + ["impl", "cpu_baseline", "e2e"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == METRIC and d["unit"] == "tokens/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference") and d["value"] > 0
