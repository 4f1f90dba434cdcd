import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than a few seconds")


def golden_names():
    """Reference inputs with the reference's golden numbers (make_golden.py); the fast__*.json files are the fast
    oracle's fixtures for the synthetic benchmark inputs (make_fast_fixtures.py)."""
    return sorted(f[:-5] for f in os.listdir(GOLDEN) if f.endswith(".json") and not f.startswith("fast__"))


def fast_fixture(case):
    with open(os.path.join(GOLDEN, f"fast__{case}.json")) as fh:
        return json.load(fh)


def fast_fixture_names():
    return sorted(f[6:-5] for f in os.listdir(GOLDEN) if f.startswith("fast__") and f.endswith(".json"))


def load_golden(name):
    from valence_b200 import inputs
    with open(os.path.join(GOLDEN, name + ".json")) as fh:
        d = json.load(fh)
    return inputs.ValenceInput.from_json(d["input"]), d["golden"]


@pytest.fixture
def write_input(tmp_path):
    """Write a ValenceInput (or a golden fixture by name) as a VALENCE input file."""
    from valence_b200 import inputs

    def _w(obj, fname="case.inp"):
        gold = None
        if isinstance(obj, str):
            obj, gold = load_golden(obj)
        p = tmp_path / fname
        p.write_text(inputs.write(obj))
        return str(p), gold

    return _w
