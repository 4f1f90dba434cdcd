import sys, numpy as np, tempfile, os
sys.path.insert(0, '/root/repo')
from valence_b200 import api, inputs
for n, rot in ((32, True), (4, False), (128, False)):
    inp = inputs.water_cluster(n, tol=(10, 20, 10), rotate=rot)
    p = tempfile.mktemp(suffix='.inp'); open(p, 'w').write(inputs.write(inp))
    eng = api.Engine(p); r0 = eng.energy()
    for sh in ([1.37, -2.11, 0.59], [13.7, -21.1, 5.9], [0.25, 0.5, -0.125]):
        x = np.array(inp.coords, dtype=float) + np.array(sh)
        r1 = eng.energy(x.flatten())
        print("TI", n, sh, "dE %.2e" % (r1["energy"] - r0["energy"]), "dEnuc %.2e" % (r1["enucrep"] - r0["enucrep"]), flush=True)
    eng.close(); os.unlink(p)
