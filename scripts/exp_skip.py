"""Experiment: cost breakdown of k_ptile from builds with parts compiled out (-DVB_EXP_SKIP=mask, see vb_ptile.cuh).
python scripts/exp_skip.py n suffix [suffix...]   -- one subprocess per library variant."""
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
print("RESULT", json.dumps({"E": repr(r["energy"]), "t_tiles_ms": r["t_tiles_ms"], "t_diag_ms": r["t_diag_ms"], "primq": r["n_prim_quartets"], "gflop": r["flops_model"] / 1e9}))
e.close(); os.unlink(p)
""" % ROOT
n = sys.argv[1]
for suf in sys.argv[2:]:
    env = dict(os.environ)
    if suf != "base":
        env["VB_LIB_SUFFIX"] = suf
    out = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True)
    print("==", n, suf, flush=True)
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("RESULT")), flush=True)
    if out.returncode != 0:
        print(out.stderr[-1500:])
