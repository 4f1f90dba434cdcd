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


def test_committed_gpu_arm_lines_carry_the_contract_keys():
    """The lines bench.py printed on 1/2/4/8 B200 (profiles/r2_bench_n*_H2O256.json): every key of the bench contract, the same metric,
    unit and workload at every N, strong scaling, one energy to the last digit."""
    lines = {}
    for n in (1, 2, 4, 8):
        with open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}_H2O256.json")) as fh:
            lines[n] = json.loads(fh.read().strip().splitlines()[-1])
    for n, d in lines.items():
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                  "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["metric"] == "contracted_shell_quartets_per_s" and d["unit"] == "shell quartets/s"
        assert d["dtype"] == "f64" and d["scaling"] == "strong" and d["higher_is_better"] is True and d["warmup"] >= 3
        assert d["config"]["workload"] == lines[1]["config"]["workload"] and "model" not in d["config"]
        assert d["gpu_launches"] > 0
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
        assert 0 < d["e2e"]["value"] < d["value"]                    # the end-to-end rate includes the copies and the host build
        r = d["roofline"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["energy_hartree"] == lines[1]["energy_hartree"]
    assert "cpu_baseline" in lines[1] and lines[1]["cpu_baseline"]["kind"] == "port"
    assert lines[8]["value"] > 6.5 * lines[1]["value"]                # > 0.81 on the device clock (measured 0.90)
