"""GPU energies of the water clusters against the fast-oracle fixtures (tests/golden/fast__w*.json)."""
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
    print("RESULT n", n, "E", repr(r["energy"]), "dE vs fast oracle %%+.2e" %% (r["energy"] - fx["energy"]), "e1 %%r e2 %%r wfnorm %%r" %% (r["e1"], r["e2"], r["wfnorm"]), "fx num %%r wf %%r" %% (fx["numerator"], fx["wfnorm"]), "minpiv", r["min_pivot_ratio"], flush=True)
""" % (ROOT, ROOT)
for s in [{}, {"VB_PRIM_TAU": "1e-24"}]:
    env = dict(os.environ); env.update(s)
    ns = sys.argv[1:] if "VB_FAST_MIN_N" not in s else [a for a in sys.argv[1:] if int(a) <= 16]
    out = subprocess.run([sys.executable, "-c", CHILD] + ns, env=env, capture_output=True, text=True)
    print("==", s, flush=True)
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("RESULT")), flush=True)
    if out.returncode != 0:
        print(out.stderr[-1500:])
