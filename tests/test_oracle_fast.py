"""The fast CPU oracle (oracle/vo_fast.c: inverse-form cofactors + AO-block integrals, same task list and screens)
against the literal restatement of the reference (oracle/valence_oracle.c) wherever the literal one finishes in
seconds, and against the reference's own golden energies.  The fast oracle is what pins the benchmark-sized inputs
(tests/golden/fast__*.json, checked against the GPU in tests/test_gpu_parity.py)."""
import math
import os

import numpy as np
import pytest

from conftest import fast_fixture, fast_fixture_names, golden_names, load_golden
from oracle.oracle import Oracle, lib
from valence_b200 import inputs

ROBUST = ("schwarz_erep", "schwarz_exch", "shortcut", "int2e_calls", "shell_quartets_2e")
SLOW = {"examples__nme3", "testing__ethane2", "testing__ethane", "examples__c3h8", "testing__c4h8.sccc.2", "testing__c8h16.vshf.2"}


def _closed_or_open_shell():
    out = []
    for n in golden_names():
        if n in SLOW:
            continue
        inp, gold = load_golden(n)
        if inp.npair == 0 and gold.get("guess_energy") is not None:
            out.append(n)
    return out


def test_boys_function_of_the_fast_path_is_exact():
    L = lib()
    import ctypes as C
    F = (C.c_double * 20)()
    L.vo_boys.argtypes = [C.c_int, C.c_double, C.c_void_p]
    rng = np.random.default_rng(7)
    worst = 0.0
    for T in list(rng.uniform(0, 60, 400)) + [0.0, 1e-9, 0.05, 34.99, 35.0, 45.95, 46.0, 46.05, 200.0]:
        for m in (0, 1, 2, 4, 8):
            L.vo_boys(m, float(T), F)
            ref = F[m]
            got = L.vo_fast_boys(m, float(T))
            # beyond the table (T >= 46) the exp(-T) terms of the upward recursion are dropped: < 1e-20 absolute
            worst = max(worst, (abs(got - ref) - 1e-21) / max(abs(ref), 1e-300))
    assert worst < 5e-15


@pytest.mark.parametrize("name", _closed_or_open_shell())
def test_fast_equals_literal_on_reference_inputs(name, write_input):
    path, gold = write_input(name)
    inp, _ = load_golden(name)
    o = Oracle(path)
    rf = o.fast_guess_energy()
    rl = o.guess_energy()
    o.close()
    assert rf["enucrep"] == rl["enucrep"]
    # the literal determinants skip Givens rotations below dtol (givens.F90:245): only inputs with a tight dtol are
    # exact on the literal side ("Tolerances")
    tol = 1e-11 if inp.ntol_d >= 16 else 1e-9
    assert abs(rf["energy"] - rl["energy"]) < tol, (rf["energy"], rl["energy"])
    assert abs(rf["energy"] - gold["guess_energy"]) < 1e-9
    for k in ROBUST:
        assert rf["counters"][k] == rl["counters"][k], k
    if inp.ntol_i <= 14:     # at itol = 1e-20 the value screen compares rounding noise of symmetry-forbidden integrals
        for k in ("value_erep", "value_exch"):
            assert rf["counters"][k] == rl["counters"][k], k


def _spin_coupled():
    out = []
    for n in golden_names():
        if n in SLOW or n in ("testing__f2-scval-p2", "testing__n2.sc4val-b.p2"):
            continue
        inp, gold = load_golden(n)
        if inp.npair > 0 and gold.get("guess_energy") is not None:
            out.append(n)
    return out


@pytest.mark.parametrize("name", _spin_coupled())
def test_fast_equals_literal_on_spin_coupled_reference_inputs(name, write_input):
    """Spin-coupled wavefunctions through the fast path (f_cofactors_sc: one inverse pair per determinant pair of density_sc /
    dbra / dket, valence.F90:1612-1870) against the literal restatement and the reference's golden energies: one and two pairs,
    one and two spin couplings."""
    path, gold = write_input(name)
    inp, _ = load_golden(name)
    o = Oracle(path)
    rl = o.guess_energy()
    try:
        rf = o.fast_guess_energy()
    except RuntimeError:
        o.close()
        pytest.skip("singular spin block (symmetry-orthogonal orbitals): outside the fast path, the literal oracle covers it")
    o.close()
    tol = 1e-11 if inp.ntol_d >= 16 else 1e-9
    assert abs(rf["energy"] - rl["energy"]) < tol, (rf["energy"], rl["energy"])
    assert abs(rf["wfnorm"] / rl["wfnorm"] - 1.0) < 1e-11
    assert abs(rf["energy"] - gold["guess_energy"]) < 1e-9
    for k in ROBUST:
        assert rf["counters"][k] == rl["counters"][k], k


@pytest.mark.parametrize("n,sc", [(2, 1), (2, 2)])
def test_fast_equals_literal_on_spin_coupled_water_clusters(n, sc, write_input):
    """Config 5's spin-coupled variant at the sizes the literal oracle reaches: 4 and 256 determinant pairs."""
    path, _ = write_input(inputs.water_cluster(n, tol=(10, 20, 10), sc_molecules=sc))
    o = Oracle(path)
    rf = o.fast_guess_energy()
    rl = o.guess_energy()
    o.close()
    assert abs(rf["energy"] - rl["energy"]) < 1e-11, (rf["energy"], rl["energy"])
    assert abs(rf["wfnorm"] / rl["wfnorm"] - 1.0) < 1e-11
    for k in ROBUST + ("value_erep", "value_exch"):
        assert rf["counters"][k] == rl["counters"][k], k


@pytest.mark.parametrize("n,rot", [(2, False), (3, True)])
def test_fast_equals_literal_on_water_clusters(n, rot, write_input):
    path, _ = write_input(inputs.water_cluster(n, tol=(10, 20, 10), rotate=rot))
    o = Oracle(path)
    rf = o.fast_guess_energy()
    rl = o.guess_energy()
    o.close()
    assert abs(rf["energy"] - rl["energy"]) < 1e-11
    assert abs(rf["wfnorm"] / rl["wfnorm"] - 1.0) < 1e-12
    for k in ROBUST + ("value_erep", "value_exch"):
        assert rf["counters"][k] == rl["counters"][k], k


def test_fast_is_reproducible_for_any_thread_count(write_input):
    path, _ = write_input(inputs.water_cluster(2, tol=(10, 20, 10)))
    e = []
    for t in (1, 3, 8):          # a fresh context each time: every energy call re-normalises the weights in place
        o = Oracle(path)
        e.append(o.fast_guess_energy(nthreads=t)["energy"])
        o.close()
    assert e[0] == e[1] == e[2]


@pytest.mark.parametrize("name,iorb", [("examples__h2o", 1), ("examples__h2o", 4), ("examples__cu+.3d94s1", 5),
                                       ("examples__lih.SDVB", 2), ("examples__cu+.3d94s1", 1)])
def test_fast_first_order_matrices_equal_literal(name, iorb, write_input):
    """ham / ovl of first_order_opt (valence.F90:527-764), incl. d shells, unpaired electrons and the spin-average pass."""
    path, _ = write_input(name)
    o = Oracle(path)
    Hf, Sf = o.fast_first_order(iorb)
    H, S, _ = o.first_order(iorb)
    o.close()
    assert np.abs(Hf - H).max() < 1e-10 and np.abs(Sf - S).max() < 1e-12


def test_fast_first_order_on_a_cluster(write_input):
    path, _ = write_input(inputs.water_cluster(2, tol=(10, 20, 10)))
    o = Oracle(path)
    Hf, Sf = o.fast_first_order(7)
    H, S, _ = o.first_order(7)
    o.close()
    assert np.abs(Hf - H).max() < 1e-10 and np.abs(Sf - S).max() < 1e-12


def test_fast_refuses_what_it_does_not_cover(write_input):
    path, _ = write_input("examples__h2o.SC")     # spin-coupled pair: energies are covered, first_order_opt matrices are not
    o = Oracle(path)
    o.fast_guess_energy()
    with pytest.raises(RuntimeError):
        o.fast_first_order(1)
    o.close()


@pytest.mark.parametrize("case", fast_fixture_names())
def test_fixture_inputs_are_the_generators_output(case):
    """The committed fast-oracle fixtures belong to the inputs the generators produce today."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_fast_fixtures as m
    fx = fast_fixture(case)
    base = m.FO_CASES[case][0] if case in m.FO_CASES else case
    assert m.digest(m.input_text(base)) == fx["input_sha256"]


def test_small_fixture_is_reproduced_live(write_input):
    fx = fast_fixture("w8")
    path, _ = write_input(inputs.water_cluster(8, tol=(10, 20, 10)))
    o = Oracle(path)
    r = o.fast_guess_energy()
    o.close()
    assert abs(r["energy"] - fx["energy"]) < 1e-11
    assert r["counters"] == fx["counters"]
