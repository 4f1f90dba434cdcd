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


def _input_tokens(inp):
    """The token line tests/host/test_input_host.cpp prints, from the Python data model."""
    t = [inp.natom, inp.natom_t, inp.npair, inp.nunpd, inp.ndocc, inp.totlen, inp.xpmax, inp.nspinc, inp.num_sh, inp.num_pr, inp.nang,
         inp.ndf, inp.nset, inp.nxorb, inp.mxctr, inp.ntol_c, inp.ntol_d, inp.ntol_i, inp.ntol_e_min, inp.ntol_e_max, inp.max_iter,
         inp.ptbnmax, inp.feather, 2 * len(inp.orbset)]
    t += [v for pr in inp.orbset for v in pr]
    t += list(inp.atom_t)
    t += [x for c in inp.coords for x in c]
    for ty in inp.types:
        t += [ty.charge, len(ty.shells)]
        for s in ty.shells:
            t += [s.l, len(s.exps)]
            for e, c in zip(s.exps, s.coefs):
                t += [e, c]
    t += [len(inp.coeff_sc)] + list(inp.coeff_sc)
    flat = [v for cpl in inp.pair_sc for pr in cpl for v in pr]
    t += [len(flat)] + flat
    t += [2 * len(inp.xorb)] + [v for pr in inp.xorb for v in pr]
    t += [len(inp.orbitals)]
    for o in inp.orbitals:
        t += [len(o.atoms)] + list(o.atoms) + [len(o.terms)]
        for xp, c in o.terms:
            t += [xp, c]
    nelec = 2 * inp.npair + 2 * inp.ndocc + inp.nunpd
    t += [nelec, 2 * inp.npair + inp.ndocc + inp.nunpd + inp.ndf, inp.npair + inp.nunpd + inp.ndocc, inp.npair + inp.ndocc,
          2 * inp.npair + inp.nunpd]
    return t


def test_input_reader_reads_every_reference_input_like_the_data_model(tmp_path):
    """vb_input.cpp on the CPU: all reference inputs of tests/golden (examples/ and testing/ of the reference, written back by
    inputs.write) and the synthetic generators parse to exactly the values of the Python data model, value by value; truncated
    files are refused with an error, not read short (xm_module.F90:41-287 stops on a short read)."""
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import golden_names, load_golden
    from valence_b200 import inputs
    cases = [(n, load_golden(n)[0]) for n in golden_names()]
    cases += [("w3sc", inputs.water_cluster(3, sc_molecules=1)), ("lif", inputs.lif_cluster(2, 2, 2)), ("c4h10", inputs.alkane(4))]
    paths = []
    for n, inp in cases:
        p = tmp_path / (n + ".inp")
        p.write_text(inputs.write(inp))
        paths.append(str(p))
    text = inputs.write(cases[0][1])
    for k, frac in enumerate((0.3, 0.6, 0.9)):
        p = tmp_path / f"cut{k}.inp"
        p.write_text(text[:int(len(text) * frac)])
        paths.append(str(p))
    exe = tmp_path / "test_input"
    csrc = os.path.join(ROOT, "valence_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", csrc, "-o", str(exe), os.path.join(ROOT, "tests", "host", "test_input_host.cpp"),
                           os.path.join(csrc, "vb_input.cpp")])
    out = subprocess.run([str(exe)] + paths, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr
    lines = out.stdout.splitlines()
    assert len(lines) == len(paths)
    for (n, inp), line in zip(cases, lines):
        tok = line.split()
        assert tok[0] == "OK", (n, line[:200])
        want = _input_tokens(inp)
        assert len(tok) - 1 == len(want), (n, len(tok) - 1, len(want))
        for k, (a, b) in enumerate(zip(tok[1:], want)):
            assert float(a) == float(b), (n, k, a, b)
    for line in lines[len(cases):]:
        assert line.startswith("ERROR"), line[:200]


def test_packed_cofactor_sets_against_determinants_of_minors(tmp_path):
    """vb_cofactor.cpp on the CPU: for closed-shell, open-shell and spin-coupled wavefunctions (up to 576 determinant pairs, three
    couplings) the packed data of every determinant pair reproduces det M and its first and second derivatives with respect to the
    matrix elements -- LU determinants of the explicit minors -- to 1e-13 of the block's scale, including blocks with a one- or
    two-dimensional null space (where the reference's Givens determinants are finite and an inverse is not) and rank deficit 3
    (everything vanishes); weights c_i c_j; the one-electron numerator and the norm; the single-coupling-pair mode of spin_opt."""
    exe = tmp_path / "test_cofactor"
    csrc = os.path.join(ROOT, "valence_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", csrc, "-o", str(exe), os.path.join(ROOT, "tests", "host", "test_cofactor_host.cpp"),
                           os.path.join(csrc, "vb_cofactor.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout[-3000:] + out.stderr


def test_launcher_detection_of_the_nccl_layer(tmp_path):
    """vb_nccl.cpp on the CPU: rank / size / local rank from torchrun, Open MPI, PMI and Slurm variables, the single-process
    default, the rendezvous key (VB_NCCL_KEY | MASTER_PORT | parent pid), refusal of an impossible rank."""
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = tmp_path / "test_launch_env"
    csrc = os.path.join(ROOT, "valence_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", csrc, "-I", os.path.join(cuda, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "test_launch_env_host.cpp"), os.path.join(csrc, "vb_nccl.cpp"),
                           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-ldl", "-Wl,-rpath," + os.path.join(cuda, "lib64")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr
