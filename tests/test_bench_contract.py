"""bench.py's contract on CPU: the reference arm prints one JSON line with the required keys (small workload)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "er-50k-1m-d128", "--steps", "1", "--warmup", "1", "--cpu-iters", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "edges/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["config"]["workload"] == "er-50k-1m-d128"
    assert j["product_library_loaded"] is False          # the CPU arm builds its graph with the oracle's own ingest


def test_workload_generators_are_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    w = dict(bench.WORKLOADS["er-50k-1m-d128"])
    u1, v1 = bench.gen_pairs(w)
    u2, v2 = bench.gen_pairs(w)
    assert (u1 == u2).all() and (v1 == v2).all() and (u1 != v1).all() and len(u1) > 0.99 * w["e"]
    assert bench.spmm_bytes(10, 100, 256) == 100 * (8 + 1024) + 8 * 11 + 4 * 10 * 256
