"""Ad-hoc GPU parity sweep: GPU engine vs CPU oracle on golden fixtures + small water clusters."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import inputs, api
from oracle.oracle import Oracle

def run(name, inp, golden=None):
    p = tempfile.mktemp(prefix="vb_", suffix=".inp")
    open(p, "w").write(inputs.write(inp))
    try:
        t = time.time(); o = Oracle(p); ro = o.guess_energy(); to = time.time() - t; o.close()
        t = time.time(); e = api.Engine(p); rg = e.energy(); tg = time.time() - t
        rg2 = e.energy(); e.close()
    except Exception as ex:
        print(f"{name:24s} FAILED: {ex}"); return
    finally:
        os.unlink(p)
    co, cg = ro["counters"], rg["counters"]
    cnt_ok = all(co[k] == cg[k] for k in ("schwarz_erep", "schwarz_exch", "value_erep", "value_exch", "int2e_calls", "shell_quartets_2e", "shortcut"))
    g = f" dGold {rg['energy']-golden:+.1e}" if golden is not None else ""
    print(f"{name:24s} E_gpu {rg['energy']:.12f} dE(gpu-oracle) {rg['energy']-ro['energy']:+.2e}{g} dNuc {rg['enucrep']-ro['enucrep']:+.1e} "
          f"dNorm {rg['wfnorm']/ro['wfnorm']-1:+.1e} rerun {rg2['energy']-rg['energy']:+.1e} counts {'OK' if cnt_ok else 'DIFF'} "
          f"tiles {rg['n_tiles']} t_tiles {rg['t_tiles_ms']:.2f}ms t_gpu {tg:.2f}s t_cpu {to:.2f}s")
    if not cnt_ok:
        print("   oracle", {k: co[k] for k in cg if k in co}); print("   gpu   ", cg)

names = sys.argv[1:] or ["examples__h", "examples__he", "examples__li", "examples__be", "examples__be.DBF", "examples__f-", "examples__h2o",
                         "examples__ch4", "examples__lih.VSHF", "examples__lih.SDVB", "examples__n2.VSHF", "examples__c2h6", "examples__c3h8",
                         "examples__cu+.3d10", "examples__cu+.3d94s1", "examples__fe2+", "testing__b", "testing__li-", "testing__be+ndf",
                         "water2", "water3", "water4r"]
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for n in names:
    if n.startswith("water"):
        k = int(n[5:].rstrip("r")); inp = inputs.water_cluster(k, rotate=n.endswith("r")); run(n, inp)
    else:
        d = json.load(open(os.path.join(gold, n + ".json")))
        run(n, inputs.ValenceInput.from_json(d["input"]), d["golden"]["guess_energy"])
