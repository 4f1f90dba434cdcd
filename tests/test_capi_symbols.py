"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/valence_b200.h declares.  No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "valence_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)))


def test_header_declares_reference_entry_points():
    names = declared_symbols()
    for ref in ("valence_api_initialize_", "valence_api_calculate_energy_", "valence_api_finalize_",
                "init_", "getn_", "calcsurface_", "finalize_"):
        assert ref in names


def test_library_exports_every_declared_symbol():
    from valence_b200 import build
    lib = ctypes.CDLL(build.build())
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/valence_b200.h but not exported"


def test_fortran_binding_source_matches_header():
    """fortran/valence_b200.f90 (the ISO_C_BINDING module a Fortran host compiles) binds only symbols the header declares,
    and the result type lists the same members in the same order as the C struct."""
    f90 = open(os.path.join(ROOT, "fortran", "valence_b200.f90")).read()
    bound = set(re.findall(r'bind\(c,\s*name="([a-z_0-9]+)"\)', f90))
    declared = set(declared_symbols())
    assert bound and bound <= declared, sorted(bound - declared)
    for must in ("vb_engine_create", "vb_engine_energy", "vb_engine_first_order", "vb_engine_attach_comm", "vb_engine_run"):
        assert must in bound
    hdr = open(os.path.join(ROOT, "include", "valence_b200.h")).read()
    cstruct = hdr[hdr.index("typedef struct vb_energy_result {"):hdr.index("} vb_energy_result;")]
    cstruct = re.sub(r"/\*.*?\*/", "", cstruct, flags=re.S)
    c_members = [m for decl in re.findall(r"(?:double|long long|int)\s+([^;]+);", cstruct) for m in re.split(r",\s*", re.sub(r"\[.*?\]", "", decl).strip())]
    ftype = f90[f90.index("type, bind(c) :: vb_energy_result"):f90.index("end type vb_energy_result")]
    ftype = re.sub(r"!.*", "", ftype)
    f_members = [m for decl in re.findall(r"::\s*(.+)", ftype)[1:] for m in re.split(r",\s*", re.sub(r"\(.*?\)", "", decl).strip())]
    assert [m.strip() for m in c_members] == [m.strip() for m in f_members]


def test_engine_fails_loudly_without_gpu(write_input):
    """No CPU fallback: creating an engine without a CUDA device must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from valence_b200 import api
    path, _ = write_input("examples__h2o")
    with pytest.raises(RuntimeError):
        api.Engine(path)


def test_engine_reports_input_errors(tmp_path):
    from valence_b200 import api
    with pytest.raises(RuntimeError):
        api.Engine(str(tmp_path / "does_not_exist.inp"))


def test_orbitals_and_nelecwfn_files_follow_xm_output_formats(write_input, tmp_path, monkeypatch):
    """xm_output (xm_module.F90:464-577): `orbitals` in formats 1/2/5/7 and, for several spin couplings, `nelecwfn`
    in formats 3/4 -- the files vtools and the NITROGEN loop read back.  Expected strings are what the reference
    wrote for this very case (tail of testing/test_cases/be-scv3s+2sc)."""
    import ctypes as C
    from valence_b200 import api, inputs
    path, _ = write_input("testing__be-scv3s+2sc")
    inp = inputs.parse_file(path)
    L = api.load()
    L.vb_write_wavefunction_files.argtypes = [C.c_char_p, C.c_double, C.c_int]
    monkeypatch.chdir(tmp_path)
    assert L.vb_write_wavefunction_files(path.encode(), -14.3173525131288084, 1) == 0
    lines = (tmp_path / "orbitals").read_text().split("\n")
    assert " total energy in atomic units             -14.3173525131288084" in lines
    k = lines.index(" total energy in atomic units             -14.3173525131288084")
    assert lines[k + 1] == " converged to   0.10E-05 kCal/mol"
    assert lines[k - 1] == "" and lines[k - 2] == "" and lines[k + 2] == ""
    # first orbital: header (5x,i2,5x,10i4) and weights 4(i4,1x,f13.8) per line
    o0 = inp.orbitals[0]
    assert lines[0] == "     %2d     " % len(o0.atoms) + "".join("%4d" % a for a in o0.atoms) + "%4d" % len(o0.terms)
    assert lines[1] == "".join("%4d %13.8f" % (i, w) for i, w in o0.terms[:4])
    # every weight reads back
    got = []
    for ln in lines[:k]:
        f = ln.split()
        if len(f) >= 2 and "." in f[1]:
            got += [float(x) for x in f[1::2]]
    want = [w for o in inp.orbitals for _, w in o.terms]
    assert len(got) == len(want) and max(abs(a - b) for a, b in zip(got, want)) < 5e-9
    nl = (tmp_path / "nelecwfn").read_text().split("\n")
    assert len(nl) >= inp.nspinc + 2
    for j in range(inp.nspinc):
        assert nl[j] == " %13.8f" % inp.coeff_sc[j] + "".join("%3d%3d  " % p for p in inp.pair_sc[j])
    assert nl[inp.nspinc] == " " + "".join("%3d%3d  " % x for x in inp.xorb)
    # energy-only runs have no convergence line
    assert L.vb_write_wavefunction_files(path.encode(), -1.0, 0) == 0
    assert "converged" not in (tmp_path / "orbitals").read_text()
