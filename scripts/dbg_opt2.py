import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from valence_b200 import api, inputs
from oracle.oracle import Oracle
name = sys.argv[1]
d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden", name + ".json")))
for tol in [(20, 20, 20), (20, 12, 20), (20, 20, 10), (20, 12, 10)]:
    inp = inputs.ValenceInput.from_json(d["input"]); inp.ntol_c, inp.ntol_d, inp.ntol_i = tol
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp))
    o2 = Oracle(p); Ho, So, co = o2.first_order(1); o2.close()
    e = api.Engine(p); Hg, Sg, st = e.first_order(1); e.close()
    print(tol, "dH", np.abs(Hg-Ho).max(), "dS", np.abs(Sg-So).max())
    if tol == (20, 12, 10):
        np.set_printoptions(linewidth=200, precision=6)
        print(Ho); print(Hg); print({k: co[k] for k in ("schwarz_erep","schwarz_exch","value_erep","value_exch","int2e_calls","shortcut")}); print(st["counters"])
