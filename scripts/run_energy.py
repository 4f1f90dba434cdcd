"""One guess energy of a water cluster on the GPU: python scripts/run_energy.py n   (VB_* switches from the environment)."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from valence_b200 import inputs, api
n = int(sys.argv[1])
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    r = e.energy()
e.close(); os.unlink(p)
fx = os.path.join(ROOT, "tests", "golden", "fast__w%d.json" % n)
de = r["energy"] - json.load(open(fx))["energy"] if os.path.exists(fx) else float("nan")
print("RESULT n", n, "E", repr(r["energy"]), "dE vs fast oracle %+.2e" % de, "primq", r["n_prim_quartets"], "tiles ms %.1f" % r["t_tiles_ms"], flush=True)
