"""Experiment: primitive-quartet magnitude cut (VB_PRIM_TAU) vs energy and work, split-launch timing, class breakdown.
python scripts/exp_tau.py n   -- each setting runs in a subprocess (the env vars are read at engine creation)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, tempfile, json
sys.path.insert(0, %r)
from valence_b200 import inputs, api
n = int(sys.argv[1])
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p)
r = e.energy(); r = e.energy()
print("RESULT", json.dumps({"E": repr(r["energy"]), "t_tiles_ms": r["t_tiles_ms"], "primq": r["n_prim_quartets"], "gflop": r["flops_model"] / 1e9,
                            "tiles": r["n_tiles"], "value_erep": r["counters"]["value_erep"]}))
e.close(); os.unlink(p)
""" % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "64"
settings = [{}, {"VB_PRIM_TAU": "1e-24"}, {"VB_PRIM_TAU": "1e-20"}, {"VB_PRIM_TAU": "1e-18"}, {"VB_PRIM_TAU": "1e-16"}, {"VB_PRIM_TAU": "1e-14"},
            {"VB_SPLIT": "1", "VB_DEBUG_TIME": "1"}, {"VB_DEBUG_PQ": "1"}]
for s in settings:
    env = dict(os.environ); env.update(s)
    out = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True)
    print("==", n, s, flush=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT") or "split chunk" in l or l.startswith("class")]
    print("\n".join(lines[-24:]), flush=True)
    if out.returncode != 0:
        print(out.stderr[-2000:])
