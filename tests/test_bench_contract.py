"""bench.py contract (no GPU): the reference arm prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout                      # build chatter and library banners go to stderr
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    d = _run("--impl", "reference", "--size", "12", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["metric"] == "ipm_iterations_per_sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_on_the_sharded_workload_and_on_other_ranks():
    d = _run("--impl", "reference", "--workload", "sphere", "--size", "5", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and "sphere_packing" in d["config"]["workload"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""    # ranks other than 0 exit without work
