"""Regenerate tests/golden/*.json from the reference checkout.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Each fixture = the parsed numeric content of one reference input
(examples/* or testing/test_cases/*) plus the golden numbers the reference's
own acceptance scripts hold for it (examples/test_examples.py:63-151,
testing/testing.py:69-151).  The fixtures are data, not reference source; the
tests rebuild an input file from them with valence_b200.inputs.write().
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from valence_b200 import inputs  # noqa: E402

REF = "/root/reference"
PAT = re.compile(r"vsvb_output\(\s*([-\d.eE+]+)\s*,\s*([-\d.eE+]+)\s*,\s*(True|False)\s*,"
                 r"\s*([-\d.eE+]+)\s*,\s*\"([^\"]+)\"", re.S)


def goldens(script):
    out = {}
    for nuc, guess, conv, tot, name in PAT.findall(open(script).read()):
        out[name] = {"nuclear_repulsion": float(nuc), "guess_energy": float(guess),
                     "converges": conv == "True", "total_energy": float(tot)}
    return out


def main():
    sets = [("examples", os.path.join(REF, "examples"), goldens(os.path.join(REF, "examples/test_examples.py"))),
            ("testing", os.path.join(REF, "testing/test_cases"), goldens(os.path.join(REF, "testing/testing.py")))]
    for tag, d, gold in sets:
        for name, g in sorted(gold.items()):
            inp = inputs.parse_file(os.path.join(d, name))
            inputs.dump_json(inp, os.path.join(HERE, f"{tag}__{name}.json"),
                             name=name, source=f"{tag}/{name}", golden=g)
            print(tag, name, inp.nelec, g["guess_energy"])


if __name__ == "__main__":
    main()
