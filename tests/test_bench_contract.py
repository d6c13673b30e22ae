"""bench.py's reference arm runs on the CPU (the C oracle stands in for the Dart reference), so its side of the JSON
contract can be checked here: one line, the arm's keys, and that a rank other than 0 prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *args):
    env = dict(os.environ, B200Z_REF_UNITS="64", B200Z_CACHE=os.path.join(ROOT, ".pytest_cache", "b200z_cache"), **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True,
                       text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


def test_reference_arm_line():
    lines = run({}, "--steps", "2", "--warmup", "1")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "inflate_uncompressed_GBps" and d["unit"] == "GB/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("gzip-multimember-64KiB")


def test_reference_arm_other_ranks_are_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "1") == []
