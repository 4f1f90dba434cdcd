"""One energy evaluation of (H2O)_n for ncu captures: python scripts/ncu_tile.py n [reps]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import inputs, api
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p)
for _ in range(reps):
    r = e.energy()
print(n, r["energy"], r["t_tiles_ms"])
e.close(); os.unlink(p)
