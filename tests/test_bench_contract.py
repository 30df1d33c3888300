"""CPU: `bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) prints exactly ONE JSON line on stdout
with the keys the contract names, whatever libraries print while it runs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--no-hnsw"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("queries/sec @1Mx384-d top_k=5") and j["value"] > 0
    assert j["config"]["rows"] == 1_000_000 and j["config"]["dim"] == 384 and j["config"]["top_k"] == 5
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0
