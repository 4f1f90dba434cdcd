import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import api, inputs
from oracle.oracle import Oracle
d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden", "testing__h2-dz.json")))
for terms in ([(6, 0.7), (10, 0.7)], [(1, 0.7), (5, 0.7)], [(6, 0.7), (10, 0.7), (1, 0.2)], [(7, 0.7), (10, 0.7)], [(6, 0.7), (8, 0.7)]):
    inp = inputs.ValenceInput.from_json(d["input"]); inp.orbitals[0].terms = terms; inp.fix_counts(); inp.max_iter = 0; inp.nset = 0; inp.orbset = []
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp))
    o = Oracle(p); ro = o.guess_energy(); o.close()
    e = api.Engine(p); r = e.energy(); e.close()
    print(terms, "oracle", ro["energy"], "gpu", r["energy"], "d", r["energy"] - ro["energy"], "e2", r["e2"], "num", r["numerator"], ro["numerator"])
