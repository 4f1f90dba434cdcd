"""first_order_opt matrices: rank-one / F-matrix path against the element-by-element contraction path, and timing."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, tempfile, json, time
sys.path.insert(0, %r)
import numpy as np
from valence_b200 import inputs, api
n, iorb = int(sys.argv[1]), int(sys.argv[2])
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p)
t = time.time(); H, S, st = e.first_order(iorb); dt = time.time() - t
r = e.energy()
np.save(sys.argv[3], np.stack([H, S]))
cw = np.array([w for _, w in inputs.water_cluster(n, tol=(10, 20, 10)).orbitals[iorb - 1].terms])
print("RESULT", json.dumps({"n": n, "orb": iorb, "seconds": round(dt, 2), "kernel_ms": round(st["t_tiles_ms"], 1), "launches": st["launches"],
      "rayleigh_minus_E": float(cw @ H @ cw / (cw @ S @ cw)) + r["enucrep"] - r["energy"], "int2e": st["counters"]["int2e_calls"], "shellq": st["counters"]["shell_quartets_2e"]}))
e.close(); os.unlink(p)
""" % ROOT
import numpy as np
for n, iorb in [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]]:
    outs = []
    for s in ({}, {"VB_FO_RANK1": "0"}):
        if n > 64 and s: continue
        env = dict(os.environ); env.update(s)
        f = f"/tmp/fo_{n}_{iorb}_{len(s)}.npy"
        out = subprocess.run([sys.executable, "-c", CHILD, str(n), str(iorb), f], env=env, capture_output=True, text=True)
        print("==", n, iorb, s, "\n".join(l for l in out.stdout.splitlines() if l.startswith(("RESULT", "[fo]", "[time]"))), flush=True)
        if out.returncode != 0: print(out.stderr[-1500:])
        else: outs.append(np.load(f))
    if len(outs) == 2:
        print("   max |dH| / |H|max %.2e   max |dS| / |S|max %.2e" % (abs(outs[0][0] - outs[1][0]).max() / abs(outs[1][0]).max(), abs(outs[0][1] - outs[1][1]).max() / abs(outs[1][1]).max()), flush=True)
