"""CPU: the JSON line contract of bench.py, exercised through the arm that needs no GPU (--impl reference = the oracle
timed on host cores, on a tiny ray sample here so the whole check takes seconds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, NSFF_BENCH_CPU_SAMPLE_RAYS="32", OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("ray-samples/sec") and d["unit"] == "ray-samples/s" and d["value"] > 0
    assert d["config"]["workload"].startswith("full NSFF") and "BASELINE.json configs[1]" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["gpu_launches"] == 0
