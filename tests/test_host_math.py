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


def test_pair_table_build_is_thread_count_independent(tmp_path):
    """vb_setup.cpp on the CPU: identical tables for 1 and 5 host threads (multi-GPU ranks must derive
    the same tile list), flat primitive lists = sorted permutations, compact Boys table vs exact series."""
    import sys
    sys.path.insert(0, ROOT)
    from valence_b200 import inputs
    inp = tmp_path / "w6.inp"
    inp.write_text(inputs.write(inputs.water_cluster(6, tol=(10, 20, 10), rotate=True)))
    exe = tmp_path / "test_setup"
    csrc = os.path.join(ROOT, "valence_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", csrc, "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "test_setup_host.cpp"),
                           os.path.join(csrc, "vb_setup.cpp"), os.path.join(csrc, "vb_input.cpp")])
    out = subprocess.run([str(exe), str(inp)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_tile_list_and_rank_sharding_on_the_host(tmp_path):
    """vb_tilelist.cpp on the CPU: the tile list is exactly the Schwarz-screened set of pair-group pairs, independent
    of the host thread count, and the work items of 1/2/3/8 ranks partition it (the N > 1 data path has no other
    exchange than the final all-reduce)."""
    exe = tmp_path / "test_tilelist"
    csrc = os.path.join(ROOT, "valence_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I", csrc, "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "test_tilelist_host.cpp"), os.path.join(csrc, "vb_tilelist.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_far_field_segment_form_against_the_recurrence(tmp_path):
    """vb_far.cuh on the CPU: the closed far-field form of (ss|ss) (ps|ss) (ss|ps) (ps|ps) over collinear shell-pair
    segments equals the Obara-Saika recurrence with the general Boys function (block-relative 5e-15), the
    segment-level far test never admits a quartet with T < 40, and the integer magnitude cut keeps everything the
    exact comparison keeps."""
    exe = tmp_path / "test_far"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "test_far_host.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
