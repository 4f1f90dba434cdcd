"""Build libvalence_b200.so (and the `valence` CLI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvalence_b200" + os.environ.get("VB_LIB_SUFFIX", "") + ".so")
CLI = os.path.join(HERE, "valence")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
SOURCES = ["vb_engine.cu", "vb_setup.cpp", "vb_tilelist.cpp", "vb_cofactor.cpp", "vb_input.cpp", "vb_capi.cpp", "vb_nccl.cpp"]
HEADERS = ["vb_engine.h", "vb_setup.h", "vb_cofactor.h", "vb_input.h", "vb_eri.cuh", "vb_kernels.cuh", "vb_tile.cuh", "vb_ptile.cuh", "vb_pclass.cuh", "vb_pseg.cuh", "vb_far.cuh", "vb_nccl.h", "vb_tilelist.h",
           os.path.join("..", "..", "include", "valence_b200.h")]


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(CLI))
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS + ["vb_cli.cpp"]]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [NVCC, *FLAGS, *os.environ.get("VB_EXTRA_FLAGS", "").split(), "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libvalence_b200.so")
    if verbose:
        print(log)
    cmd = [NVCC, "-O2", "-std=c++17", "-o", CLI, os.path.join(CSRC, "vb_cli.cpp"),
           "-L" + HERE, "-lvalence_b200", "-ldl", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building the valence CLI")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
