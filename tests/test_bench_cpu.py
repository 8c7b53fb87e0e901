"""bench.py contract on a box without a GPU: the reference arm runs (the reference's own CPU functions
on the host cores) and prints ONE JSON line with the agreed keys; our arm refuses to run without CUDA
instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--workload", "config1", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "genotype_effect_updates_per_s" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["config"]["workload"] == "config1"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_has_no_cpu_path(has_gpu):
    if has_gpu:
        return
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
