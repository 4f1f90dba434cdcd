"""GPU vs oracle: spin-coupled energies and first_order_opt matrices."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from valence_b200 import inputs, api
from oracle.oracle import Oracle
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
def load(n):
    d = json.load(open(os.path.join(gold, n + ".json"))); return inputs.ValenceInput.from_json(d["input"]), d["golden"]
def wr(inp):
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp)); return p
sc = ["examples__h2.sz", "examples__h2.dz", "examples__he1s2s", "examples__be.sv", "examples__be.2SC", "examples__be2s3s.2SC", "examples__h2o.SC",
      "examples__lih.SCval", "examples__lih.exstate", "testing__lih", "testing__lih-sv", "testing__h2o-vdz-sc1", "testing__be-sc"]
for n in sc:
    inp, g = load(n); p = wr(inp)
    try:
        o = Oracle(p); ro = o.guess_energy(); o.close()
        e = api.Engine(p); t=time.time(); r = e.energy(); dt=time.time()-t; e.close()
        co, cg = ro["counters"], r["counters"]
        ok = all(co[k] == cg[k] for k in ("schwarz_erep", "schwarz_exch", "int2e_calls", "shell_quartets_2e", "shortcut"))
        print(f"{n:26s} dE(gpu-oracle) {r['energy']-ro['energy']:+.2e} dGold {r['energy']-g['guess_energy']:+.1e} dNorm {r['wfnorm']/ro['wfnorm']-1:+.1e} counts {'OK' if ok else 'DIFF'} {dt:.2f}s")
    except Exception as ex:
        print(f"{n:26s} FAILED {ex}")
    os.unlink(p)
fo = [("examples__h2o", 1), ("examples__h2o", 4), ("examples__li", 1), ("examples__li", 2), ("examples__be.2SC", 1), ("examples__be.DBF", 2),
      ("examples__cu+.3d94s1", 1), ("examples__cu+.3d94s1", 5), ("examples__ch4", 2), ("examples__h2o.SC", 3), ("examples__lih.SCval", 1)]
for n, iorb in fo:
    inp, g = load(n); p = wr(inp)
    try:
        o = Oracle(p); t=time.time(); Ho, So, co = o.first_order(iorb); to=time.time()-t; o.close()
        e = api.Engine(p); t=time.time(); Hg, Sg, st = e.first_order(iorb); tg=time.time()-t; e.close()
        print(f"{n:26s} orb {iorb} n={len(Ho)} max|dH| {np.abs(Hg-Ho).max():.2e} max|dS| {np.abs(Sg-So).max():.2e} |H|max {np.abs(Ho).max():.1f} t_gpu {tg:.2f}s t_cpu {to:.2f}s")
    except Exception as ex:
        print(f"{n:26s} orb {iorb} FAILED {ex}")
    os.unlink(p)
