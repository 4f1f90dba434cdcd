"""Experiment: lane efficiency of the class kernels (library built with -DVB_EXP_EFF, suffix _eff)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, tempfile
sys.path.insert(0, %r)
from valence_b200 import inputs, api
n = int(sys.argv[1])
p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inputs.water_cluster(n, tol=(10, 20, 10))))
e = api.Engine(p); r = e.energy(); e.close(); os.unlink(p)
""" % ROOT
env = dict(os.environ); env["VB_LIB_SUFFIX"] = "_eff"; env["VB_DEBUG_PQ"] = "1"
out = subprocess.run([sys.executable, "-c", CHILD, sys.argv[1]], env=env, capture_output=True, text=True)
print("\n".join(l for l in out.stdout.splitlines() if "lane efficiency" in l or "tasks" in l))
print(out.stderr[-800:])
