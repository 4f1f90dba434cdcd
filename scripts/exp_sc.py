"""Spin-coupled water clusters: GPU determinant-pair path (forced with VB_FAST_MIN_N=0) against the host factorisation and the literal oracle."""
import os, subprocess, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, tempfile, time
sys.path.insert(0, %r)
from valence_b200 import inputs, api
n, sc = int(sys.argv[1]), int(sys.argv[2])
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10), sc_molecules=sc)))
e = api.Engine(p); t = time.time(); r = e.energy(); dt = time.time() - t; e.close()
print("RESULT", n, sc, repr(r["energy"]), repr(r["wfnorm"]), "%%.2f s" %% dt, "launches", r["launches"], "minpiv %%.3g" %% r["min_pivot_ratio"], flush=True)
if len(sys.argv) > 3:
    from oracle.oracle import Oracle
    o = Oracle(p); ro = o.guess_energy(); o.close()
    print("ORACLE", repr(ro["energy"]), "dE %%.2e" %% (r["energy"] - ro["energy"]), all(ro["counters"][k] == r["counters"][k] for k in ("schwarz_erep", "schwarz_exch", "int2e_calls", "shell_quartets_2e", "shortcut")), flush=True)
os.unlink(p)
""" % ROOT
def run(n, sc, env, oracle=False):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD, str(n), str(sc)] + (["o"] if oracle else []), env=e, capture_output=True, text=True)
    print(env, "\n".join(l for l in out.stdout.splitlines() if l.startswith(("RESULT", "ORACLE"))), out.stderr[-300:] if out.returncode else "", flush=True)
run(2, 1, {"VB_FAST_MIN_N": "0"}, True)
run(2, 2, {"VB_FAST_MIN_N": "0"}, True)
run(3, 1, {"VB_FAST_MIN_N": "0"}, True)
for n, sc in ((8, 1), (8, 2)):
    run(n, sc, {})
    run(n, sc, {"VB_FAST_MIN_N": "0"})
run(16, 2, {})
run(32, 2, {})
run(64, 1, {})
