"""CPU: the parts of bench.py's contract that do not need a GPU - the reference arm (`--impl reference`, the CPU oracle
port) prints exactly ONE JSON line on stdout with the agreed keys, chatter goes to stderr, and under a multi-rank
launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--lr", "32", "--batch", "2", "--steps", "1", "--warmup", "0"]


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run(CMD, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = run({"NCCL_DEBUG": "VERSION"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/femasr_oracle.py" in cb["sample"]
    assert d["config"]["workload"].startswith("config 2") and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
