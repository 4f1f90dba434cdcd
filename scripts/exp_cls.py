"""Experiment: class-split pass (vb_pclass.cuh) vs the all-in-one kernel: energy, time, per-class times."""
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
                            "TF": r["flops_model"] / r["t_tiles_ms"] / 1e9, "cnt": r["counters"]["shell_quartets_2e"], "verep": r["counters"]["value_erep"]}))
e.close(); os.unlink(p)
""" % ROOT
n = sys.argv[1]
for s in [{"VB_CLASS_SPLIT": "0"}, {"VB_CLASS_SPLIT": "1", "VB_DEBUG_TIME": "1"}, {"VB_CLASS_SPLIT": "1"}] + [json.loads(a) for a in sys.argv[2:]]:
    env = dict(os.environ); env.update(s)
    out = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True)
    print("==", n, s, flush=True)
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("RESULT") or "class pass" in l), flush=True)
    if out.returncode != 0:
        print(out.stderr[-1500:])
