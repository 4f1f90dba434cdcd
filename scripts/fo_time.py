"""Energy + first_order_opt accumulators of (H2O)_n, timed: python scripts/fo_time.py n [iorb]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import inputs, api
n = int(sys.argv[1]); iorb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p)
t0 = time.perf_counter(); r = e.energy(); t1 = time.perf_counter()
H, S, st = e.first_order(iorb); t2 = time.perf_counter()
print(f"n {n} energy {r['energy']:.10f} in {t1-t0:.2f} s; first_order({iorb}) {H.shape} in {t2-t1:.2f} s, tile kernels {st['t_tiles_ms']:.0f} ms, "
      f"prim quartets {st['n_prim_quartets']:.3e}, ham[0,0] {H[0,0]:.10f} ovl[0,0] {S[0,0]:.12f}", flush=True)
# Rayleigh quotient of the current weights reproduces the energy numerator / norm: c^T H c / c^T S c + enuc = E
import numpy as np
inp = inputs.water_cluster(n, tol=(10, 20, 10))
c = np.array([w for _, w in inp.orbitals[iorb - 1].terms], dtype=float)
print("Rayleigh quotient c^T H c / c^T S c + enuc - E =", float(c @ H @ c / (c @ S @ c)) + r["enucrep"] - r["energy"], "  ham[0,0] %.6e ovl[0,0] %.6e" % (H[0, 0], S[0, 0]))
e.close(); os.unlink(p)
