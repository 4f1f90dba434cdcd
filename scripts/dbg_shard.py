import sys, tempfile, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from valence_b200 import api, inputs
p = tempfile.mktemp(suffix='.inp'); open(p,'w').write(inputs.write(inputs.water_cluster(4, tol=(10,20,10))))
eng = api.Engine(p)
L = eng.L
L.vb_engine_debug_tile_energies.restype = C.c_longlong
L.vb_engine_debug_tile_energies.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
def tiles():
    n = L.vb_engine_debug_tile_energies(eng.h, None, 0)
    a = np.zeros(n); L.vb_engine_debug_tile_energies(eng.h, a.ctypes.data, n); return a
full = eng.energy(); tf = tiles()
print('full e2', full['e2'], tf.sum())
r = eng.energy_partial(0, 2); t0 = tiles()
r = eng.energy_partial(1, 2); t1 = tiles()
print('nonzero', (t0!=0).sum(), (t1!=0).sum(), 'overlap', ((t0!=0)&(t1!=0)).sum())
d = t0 + t1 - tf
bad = np.nonzero(np.abs(d) > 1e-12)[0]
print('bad tiles', len(bad), bad[:20], d[bad[:10]], tf[bad[:10]])
print('even idx nonzero in t0', (t0[0::2]!=0).sum(), 'odd idx nonzero in t0', (t0[1::2]!=0).sum())
