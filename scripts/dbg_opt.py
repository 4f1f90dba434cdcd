import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from valence_b200 import api, inputs
from oracle.oracle import Oracle
name = sys.argv[1]
d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden", name + ".json")))
inp = inputs.ValenceInput.from_json(d["input"]); p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp))
print("golden", d["golden"])
o = Oracle(p, quiet=False); print(o.run()); 
for iorb in (1, 2):
    o2 = Oracle(p); Ho, So, _ = o2.first_order(iorb); o2.close()
    e = api.Engine(p); Hg, Sg, _ = e.first_order(iorb); e.close()
    print("orb", iorb, "dH", np.abs(Hg-Ho).max(), "dS", np.abs(Sg-So).max())
e = api.Engine(p); print(e.run(True)); e.close()
