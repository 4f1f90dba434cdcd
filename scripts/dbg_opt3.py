import sys, os, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from valence_b200 import api, inputs
name = sys.argv[1]
d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden", name + ".json")))
inp = inputs.ValenceInput.from_json(d["input"]); p = tempfile.mktemp(suffix=".inp"); open(p, "w").write(inputs.write(inp))
e = api.Engine(p); e.first_order(1); e.close()
