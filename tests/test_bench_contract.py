"""The reference arm of bench.py on the CPU (the arm the driver times beside the GPU arm): JSON-line contract, rank-0-only
behaviour under a multi-rank launch.  The GPU arm needs the device; its lines from the B200 runs are committed under profiles/ (r2_bench_*.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *argv):
    env = dict(os.environ, VB_BENCH_REF_SECONDS="0.5", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *argv], capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    out = _run({}, "--workload", "w4", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "contracted_shell_quartets_per_s" and d["unit"] == "shell quartets/s" and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "(H2O)_4" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "schwarz_ints" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--workload", "w4", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
