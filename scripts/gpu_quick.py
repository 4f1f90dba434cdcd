"""Quick GPU parity check (seconds): a few reference inputs and small clusters against the oracle."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import inputs, api
from oracle.oracle import Oracle
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
def wr(inp):
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp)); return p
cases = []
for n in ("examples__h2o", "examples__c2h6", "examples__h2o.SC", "examples__lih.SDVB", "examples__cu+.3d94s1"):
    d = json.load(open(os.path.join(gold, n + ".json"))); cases.append((n, inputs.ValenceInput.from_json(d["input"])))
cases.append(("water3rot", inputs.water_cluster(3, tol=(10, 20, 10), rotate=True)))
cases.append(("water4", inputs.water_cluster(4, tol=(10, 20, 10))))
bad = 0
for name, inp in cases:
    p = wr(inp)
    o = Oracle(p); ro = o.guess_energy(); o.close()
    e = api.Engine(p); r = e.energy(); r2 = e.energy(); e.close(); os.unlink(p)
    ok = all(ro["counters"][k] == r["counters"][k] for k in ("schwarz_erep", "schwarz_exch", "int2e_calls", "shell_quartets_2e", "shortcut"))
    de = r["energy"] - ro["energy"]
    bad += (abs(de) > 1e-10) or not ok
    print(f"{name:24s} dE {de:+.2e} rerun {r2['energy']-r['energy']:+.1e} counts {'OK' if ok else 'DIFF'}", flush=True)
print("QUICK", "FAIL" if bad else "PASS")
