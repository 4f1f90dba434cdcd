"""Experiment: library variants (VB_LIB_SUFFIX) of the class-split pass, per-class times.  python scripts/exp_var.py n suffix..."""
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
print("RESULT", json.dumps({"E": repr(r["energy"]), "t_tiles_ms": round(r["t_tiles_ms"], 1), "primq": r["n_prim_quartets"], "TF": round(r["flops_model"] / r["t_tiles_ms"] / 1e9, 2)}))
e.close(); os.unlink(p)
""" % ROOT
n = sys.argv[1]
for suf in sys.argv[2:]:
    env = dict(os.environ); env["VB_DEBUG_TIME"] = "1"
    if suf != "base":
        env["VB_LIB_SUFFIX"] = suf
    out = subprocess.run([sys.executable, "-c", CHILD, n], env=env, capture_output=True, text=True)
    print("==", n, suf, flush=True)
    ls = [l for l in out.stdout.splitlines() if l.startswith("RESULT") or "class pass" in l]
    print("\n".join(ls[-2:]), flush=True)
    if out.returncode != 0:
        print(out.stderr[-1500:])
