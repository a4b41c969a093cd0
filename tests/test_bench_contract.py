"""bench.py's reference arm runs on CPU: pin the JSON line's contract here (the GPU arm's
line is produced on the B200 box; same keys plus roofline / clocks)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_json_line():
    lines = _run("--impl", "reference", "--steps", "3", "--warmup", "1", "--n", "200000")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "healthcheck_schedule_evals_per_sec"
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["steps"] == 3
    assert d["value"] > 1e5 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 200000) < 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "u64"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    assert _run("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--n", "50000", env=env) == []
