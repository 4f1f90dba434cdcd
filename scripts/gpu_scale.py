"""Timing sweep of the GPU engine over synthetic water clusters."""
import os, sys, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import inputs, api
print("fp64 peak TFLOP/s", api.measure_fp64_peak(0), flush=True)
for n in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]:
    inp = inputs.water_cluster(n, tol=(10, 20, 10))
    p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp))
    e = api.Engine(p)
    for rep in range(2 if n <= 64 else 1):
        t = time.time(); r = e.energy(); dt = time.time() - t
    e.close(); os.unlink(p)
    tf = r["flops_model"] / (r["t_tiles_ms"] * 1e-3) / 1e12
    print(json.dumps({"n": n, "E": r["energy"], "wall_s": round(dt, 3), "t_tiles_ms": round(r["t_tiles_ms"], 2), "t_diag_ms": round(r["t_diag_ms"], 2),
                      "t_1e_ms": round(r["t_1e_ms"], 1), "t_density_ms": round(r["t_density_ms"], 1), "t_host_ms": round(r["t_host_setup_ms"], 1),
                      "tiles": r["n_tiles"], "pgs": r["n_pairgroups"], "primq": r["n_prim_quartets"],
                      "ref_quartets": r["ref_shell_quartets"], "model_TF": round(tf, 2), "gflop": round(r["flops_model"] / 1e9, 1),
                      "counters": r["counters"]}), flush=True)
