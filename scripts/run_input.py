"""One guess energy of a golden reference input on the GPU: python scripts/run_input.py examples__cu+.3d94s1 [reps]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from valence_b200 import inputs, api
d = json.load(open(os.path.join(ROOT, "tests", "golden", sys.argv[1] + ".json")))
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.ValenceInput.from_json(d["input"])))
e = api.Engine(p)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    t = time.time(); r = e.energy(); dt = time.time() - t
e.close(); os.unlink(p)
print("RESULT", sys.argv[1], repr(r["energy"]), "dE vs golden %+.2e" % (r["energy"] - d["golden"]["guess_energy"]), "wall %.1f ms" % (1e3 * dt),
      {k: round(r[k], 2) for k in ("t_1e_ms", "t_density_ms", "t_diag_ms", "t_tiles_ms", "t_host_setup_ms")}, "tiles", r["n_tiles"], "pair groups", r["n_pairgroups"], "launches", r["launches"], flush=True)
