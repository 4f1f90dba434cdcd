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
