"""Host-side check of the product's integral core (vb_eri.cuh is __host__ __device__): the
Obara-Saika VRR (unrolled s/p path and the loop-based d path), the Boys function and the HRR
folding agree with the oracle's independent McMurchie-Davidson integrals to ~1e-14."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vrr_and_boys_against_oracle(tmp_path):
    obj = tmp_path / "vo_integrals.o"
    exe = tmp_path / "test_eri"
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-c", "-o", str(obj), os.path.join(ROOT, "oracle", "vo_integrals.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "test_eri_host.cpp"),
                           str(obj), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
