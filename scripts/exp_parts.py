"""Components of the GPU energy (E1/c0, E2/c0, N1/nelec) of the water clusters against the fast-oracle fixtures."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, tempfile, json
sys.path.insert(0, %r)
from valence_b200 import inputs, api
for n in [int(a) for a in sys.argv[1:]]:
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
    e = api.Engine(p); r = e.energy(); e.close(); os.unlink(p)
    fx = json.load(open(os.path.join(%r, "tests", "golden", "fast__w%%d.json" %% n)))
    print("RESULT n", n, "E", repr(r["energy"]), "dE vs fast oracle %%+.2e" %% (r["energy"] - fx["energy"]), "primq", r["n_prim_quartets"], flush=True)
""" % (ROOT, ROOT)
variants = [{}, {"VB_PRIM_TAU": "1e-24"}, {"VB_INV_REFINE": "0"}]
for s in variants:
    env = dict(os.environ); env.update(s); env["VB_DEBUG_PARTS"] = "1"
    out = subprocess.run([sys.executable, "-c", CHILD] + sys.argv[1:], env=env, capture_output=True, text=True)
    print("==", s, flush=True)
    print("\n".join(l for l in (out.stdout + out.stderr).splitlines() if l.startswith("RESULT") or l.startswith("[parts]")), flush=True)
    if out.returncode != 0:
        print(out.stderr[-1500:])
