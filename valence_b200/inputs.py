"""VALENCE input files: record-based reader, writer and synthetic generators.

The grammar follows the reference's list-directed READ statements
(/root/reference/src/xm_module.F90:41-42, 106-108, 119, 133-163, 181-202,
212-287): every READ starts on a fresh record, may span several records, and
discards whatever is left on the last record it touched.  Trailing prose after
the last orbital is therefore ignored, and so is the 16th integer that
testing/test_cases/cu+ carries on its header record.

This module is host-side plumbing (fixtures, synthetic workloads); the product
library parses input files itself in C++ (csrc/vb_input.cpp).
"""
from __future__ import annotations

import copy
import dataclasses
import json
import re
from typing import List, Sequence, Tuple

import numpy as np

ANGS2BOHR = 1.889725987722  # /root/reference/src/tools_module.F90:11

_REPEAT = re.compile(r"^(\d+)\*(.*)$")


def _tokens(line: str) -> List[str]:
    out: List[str] = []
    for tok in line.replace(",", " ").split():
        if tok == "/":
            break
        m = _REPEAT.match(tok)
        if m:
            out.extend([m.group(2)] * int(m.group(1)))
        else:
            out.append(tok)
    return out


def _f(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


class _Records:
    """One Fortran list-directed READ per call."""

    def __init__(self, text: str):
        self.lines = text.splitlines()
        self.pos = 0

    def read(self, count) -> List[str]:
        """count is an int, or a callable(tokens_so_far) -> required total."""
        toks: List[str] = []
        need = count if isinstance(count, int) else count(toks)
        while len(toks) < need:
            if self.pos >= len(self.lines):
                raise EOFError("input ended inside a READ")
            toks.extend(_tokens(self.lines[self.pos]))
            self.pos += 1
            if not isinstance(count, int):
                need = count(toks)
        return toks[:need]


@dataclasses.dataclass
class Shell:
    l: int
    exps: List[float]
    coefs: List[float]


@dataclasses.dataclass
class AtomType:
    charge: float
    shells: List[Shell]


@dataclasses.dataclass
class Orbital:
    atoms: List[int]                      # 1-based atom indices (the OBS)
    terms: List[Tuple[int, float]]        # (AO index in OBS or <=0 for a DBF, weight)


@dataclasses.dataclass
class ValenceInput:
    natom: int
    natom_t: int
    npair: int
    nunpd: int
    ndocc: int
    totlen: int
    xpmax: int
    nspinc: int
    num_sh: int
    num_pr: int
    nang: int
    ndf: int
    nset: int
    nxorb: int
    mxctr: int
    ntol_c: int
    ntol_d: int
    ntol_i: int
    ntol_e_min: int
    ntol_e_max: int
    max_iter: int
    ptbnmax: float
    feather: float
    orbset: List[Tuple[int, int]]
    atom_t: List[int]
    coords: List[List[float]]             # Angstrom
    types: List[AtomType]
    coeff_sc: List[float]
    pair_sc: List[List[Tuple[int, int]]]  # [nspinc][npair]
    xorb: List[Tuple[int, int]]
    orbitals: List[Orbital]               # 2*npair, nunpd, ndocc, ndf in that order

    # --- derived (valence_initialize_module.F90:89-92)
    @property
    def nelec(self) -> int:
        return 2 * self.npair + 2 * self.ndocc + self.nunpd

    @property
    def norbs(self) -> int:
        return 2 * self.npair + self.ndocc + self.nunpd + self.ndf

    def to_json(self) -> dict:
        return dataclasses.asdict(self)

    @staticmethod
    def from_json(d: dict) -> "ValenceInput":
        d = copy.deepcopy(d)
        d["types"] = [AtomType(t["charge"], [Shell(**s) for s in t["shells"]]) for t in d["types"]]
        d["orbitals"] = [Orbital(o["atoms"], [tuple(t) for t in o["terms"]]) for o in d["orbitals"]]
        d["orbset"] = [tuple(x) for x in d["orbset"]]
        d["xorb"] = [tuple(x) for x in d["xorb"]]
        d["pair_sc"] = [[tuple(p) for p in c] for c in d["pair_sc"]]
        return ValenceInput(**d)

    def fix_counts(self) -> "ValenceInput":
        """Recompute the header counts from the content."""
        self.natom = len(self.atom_t)
        self.natom_t = len(self.types)
        self.num_sh = sum(len(t.shells) for t in self.types)
        self.num_pr = sum(len(s.exps) for t in self.types for s in t.shells)
        self.nang = max(s.l for t in self.types for s in t.shells)
        self.totlen = sum(len(o.terms) for o in self.orbitals)
        self.xpmax = max(len(o.terms) for o in self.orbitals)
        self.mxctr = max(len(o.atoms) for o in self.orbitals)
        self.nset = len(self.orbset)
        self.nxorb = len(self.xorb)
        return self


def parse(text: str) -> ValenceInput:
    r = _Records(text)
    h = [int(t) for t in r.read(15)]
    (natom, natom_t, npair, nunpd, ndocc, totlen, xpmax, nspinc,
     num_sh, num_pr, nang, ndf, nset, nxorb, mxctr) = h
    c = r.read(8 + 2 * nset)
    ntol_c, ntol_d, ntol_i, emin, emax, max_iter = (int(t) for t in c[:6])
    ptbnmax, feather = _f(c[6]), _f(c[7])
    orbset = [(int(c[8 + 2 * i]), int(c[9 + 2 * i])) for i in range(nset)]
    atom_t, coords = [], []
    for _ in range(natom):
        t = r.read(4)
        atom_t.append(int(t[0]))
        coords.append([_f(x) for x in t[1:4]])
    types = []
    for _ in range(natom_t):
        t = r.read(2)
        charge, nshell = _f(t[0]), int(t[1])
        shells = []
        for _ in range(nshell):
            t = r.read(2)
            l, k = int(t[0]), int(t[1])
            if k == 1:
                shells.append(Shell(l, [_f(r.read(1)[0])], [1.0]))
            else:
                ex, co = [], []
                for _ in range(k):
                    t = r.read(2)
                    ex.append(_f(t[0]))
                    co.append(_f(t[1]))
                shells.append(Shell(l, ex, co))
        types.append(AtomType(charge, shells))
    coeff_sc, pair_sc = [1.0], []
    if npair > 0:
        if nspinc == 1:
            t = r.read(2 * npair)
            pair_sc = [[(int(t[2 * i]), int(t[2 * i + 1])) for i in range(npair)]]
        elif nspinc > 1:
            t = r.read(nspinc * (1 + 2 * npair))
            coeff_sc, k = [], 0
            for _ in range(nspinc):
                coeff_sc.append(_f(t[k]))
                pair_sc.append([(int(t[k + 1 + 2 * i]), int(t[k + 2 + 2 * i])) for i in range(npair)])
                k += 1 + 2 * npair
    xorb = []
    if nxorb > 0:
        t = r.read(2 * nxorb)
        xorb = [(int(t[2 * i]), int(t[2 * i + 1])) for i in range(nxorb)]
    orbitals = []
    for _ in range(2 * npair + nunpd + ndocc + ndf):
        t = r.read(lambda toks: 1 if not toks else int(toks[0]) + 2)
        atnum = int(t[0])
        atoms = [int(x) for x in t[1:1 + atnum]]
        n = int(t[1 + atnum])
        t = r.read(2 * n)
        orbitals.append(Orbital(atoms, [(int(t[2 * i]), _f(t[2 * i + 1])) for i in range(n)]))
    return ValenceInput(natom, natom_t, npair, nunpd, ndocc, totlen, xpmax, nspinc, num_sh,
                        num_pr, nang, ndf, nset, nxorb, mxctr, ntol_c, ntol_d, ntol_i, emin, emax,
                        max_iter, ptbnmax, feather, orbset, atom_t, coords, types, coeff_sc,
                        pair_sc, xorb, orbitals)


def parse_file(path: str) -> ValenceInput:
    with open(path) as fh:
        return parse(fh.read())


def write(inp: ValenceInput) -> str:
    """Emit an input file the reference parser (and ours) reads back identically.
    Reals are written with 17 significant digits so the round trip is exact."""
    g = lambda x: repr(float(x))
    L: List[str] = []
    L.append(" ".join(str(v) for v in (
        inp.natom, inp.natom_t, inp.npair, inp.nunpd, inp.ndocc, inp.totlen, inp.xpmax,
        inp.nspinc, inp.num_sh, inp.num_pr, inp.nang, inp.ndf, inp.nset, inp.nxorb, inp.mxctr)))
    L.append("")
    ctl = [inp.ntol_c, inp.ntol_d, inp.ntol_i, inp.ntol_e_min, inp.ntol_e_max, inp.max_iter]
    L.append(" ".join(str(v) for v in ctl) + f" {g(inp.ptbnmax)} {g(inp.feather)} "
             + " ".join(f"{a} {b}" for a, b in inp.orbset))
    L.append("")
    for t, xyz in zip(inp.atom_t, inp.coords):
        L.append(f"{t} " + " ".join(g(x) for x in xyz))
    L.append("")
    for ty in inp.types:
        L.append(f"{g(ty.charge)} {len(ty.shells)}")
        for sh in ty.shells:
            L.append(f"{sh.l} {len(sh.exps)}")
            if len(sh.exps) == 1:
                L.append(f"  {g(sh.exps[0])}")
            else:
                for e, c in zip(sh.exps, sh.coefs):
                    L.append(f"  {g(e)} {g(c)}")
        L.append("")
    if inp.npair > 0:
        if inp.nspinc == 1:
            L.append(" ".join(f"{a} {b}" for a, b in inp.pair_sc[0]))
        else:
            for w, pairs in zip(inp.coeff_sc, inp.pair_sc):
                L.append(f"{g(w)} " + " ".join(f"{a} {b}" for a, b in pairs))
        L.append("")
    if inp.nxorb > 0:
        L.append(" ".join(f"{a} {b}" for a, b in inp.xorb))
        L.append("")
    for o in inp.orbitals:
        L.append(f"{len(o.atoms)} " + " ".join(str(a) for a in o.atoms) + f" {len(o.terms)}")
        for i in range(0, len(o.terms), 4):
            L.append(" ".join(f"{x} {g(c)}" for x, c in o.terms[i:i + 4]))
    L.append("")
    return "\n".join(L) + "\n"


# --------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------

# 6-31G water monomer: geometry (Angstrom), basis and the five optimised DOCC
# orbitals are the numeric content of the reference's examples/h2o input
# (/root/reference/examples/h2o:5-7, 9-37, 39-52), restated here as data.
_H2O_XYZ = [[0.0, 0.0, 0.1068310], [0.0, 0.7851780, -0.4273240], [0.0, -0.7851780, -0.4273240]]
_O_631G = AtomType(8.0, [
    Shell(0, [5484.6717, 825.23495, 188.04696, 52.9645, 16.89757, 5.7996353],
          [0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.470193, 0.3585209]),
    Shell(0, [15.539616, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.130767]),
    Shell(0, [0.2700058], [1.0]),
    Shell(1, [15.539616, 3.5999336, 1.0137618], [0.0708743, 0.3397528, 0.7271586]),
    Shell(1, [0.2700058], [1.0]),
])
_H_631G = AtomType(1.0, [
    Shell(0, [18.731137, 2.8253937, 0.6401217], [0.0334946, 0.23472695, 0.81375733]),
    Shell(0, [0.1612778], [1.0]),
])
_H2O_ORBS = [
    ([1, 2, 3], [(2, -0.08027660), (5, -0.29134366), (6, 0.43583969), (8, -0.15158844),
                 (9, 0.31586994), (10, -0.30665685), (11, -0.11520330), (13, 0.01734157)]),
    ([1, 2, 3], [(2, -0.08027769), (5, 0.29134376), (6, 0.43583975), (8, 0.15158846),
                 (9, 0.31586958), (11, 0.01733658), (12, -0.30665770), (13, -0.11520065)]),
    ([1, 2, 3], [(2, 0.44361598), (3, 0.52870700), (6, 0.32548255), (9, 0.25913774),
                 (11, -0.05869114), (13, -0.05868874)]),
    ([1], [(4, 0.64018474), (7, 0.51154856)]),
    ([1, 2, 3], [(1, 0.99271066), (2, 0.02865719), (6, 0.00306807), (9, 0.00582145),
                 (11, 0.00109706), (13, 0.00109709)]),
]

_GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 16: (4, 2, 2),
          32: (4, 4, 2), 64: (4, 4, 4), 128: (8, 4, 4), 256: (8, 8, 4)}


def _rot_p(R: np.ndarray, terms: Sequence[Tuple[int, float]]) -> List[Tuple[int, float]]:
    """Rotate the p-type weights of one water orbital (O p shells are OBS AOs
    4-6 and 7-9) with the molecule; s weights are invariant."""
    w = dict(terms)
    for base in (4, 7):
        v = np.array([w.get(base + k, 0.0) for k in range(3)])
        if np.any(v != 0.0):
            v = R @ v
            for k in range(3):
                w[base + k] = float(v[k])
    return sorted((k, c) for k, c in w.items() if c != 0.0)


def water_cluster(n: int, spacing: float = 3.10, tol: Tuple[int, int, int] = (10, 10, 10),
                  rotate: bool = False, seed: int = 20261017, sc_molecules: int = 0) -> ValenceInput:
    """(H2O)_n on a simple-cubic grid, 6-31G, five DOCC orbitals per monomer
    (SURVEY.md section 8d, config 5).  sc_molecules>0 converts the two OH-bond
    orbitals of the first molecules into spin-coupled pairs by the README
    recipe (duplicate, +-0.1 on the two largest weights), one Rumer coupling."""
    if n in _GRIDS:
        gx, gy, gz = _GRIDS[n]
    else:
        gx, gy, gz = n, 1, 1
    rng = np.random.default_rng(seed)
    atom_t, coords = [], []
    docc: List[Orbital] = []
    sc: List[Orbital] = []
    m = 0
    for ix in range(gx):
        for iy in range(gy):
            for iz in range(gz):
                if m >= n:
                    break
                org = np.array([ix, iy, iz], dtype=float) * spacing
                R = np.eye(3)
                if rotate:
                    q = rng.normal(size=4)
                    q /= np.linalg.norm(q)
                    a, b, c, d = q
                    R = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                                  [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                                  [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
                base = 3 * m
                for k, xyz in enumerate(_H2O_XYZ):
                    atom_t.append(1 if k == 0 else 2)
                    coords.append([float(v) for v in (R @ np.array(xyz) + org)])
                for io, (atoms, terms) in enumerate(_H2O_ORBS):
                    tt = _rot_p(R, terms) if rotate else list(terms)
                    orb = Orbital([base + a for a in atoms], tt)
                    if m < sc_molecules and io < 2:
                        big = sorted(range(len(tt)), key=lambda i: -abs(tt[i][1]))[:2]
                        t1, t2 = list(tt), list(tt)
                        t1[big[0]] = (tt[big[0]][0], tt[big[0]][1] + 0.1)
                        t1[big[1]] = (tt[big[1]][0], tt[big[1]][1] - 0.1)
                        t2[big[0]] = (tt[big[0]][0], tt[big[0]][1] - 0.1)
                        t2[big[1]] = (tt[big[1]][0], tt[big[1]][1] + 0.1)
                        sc.append(Orbital(orb.atoms, t1))
                        sc.append(Orbital(orb.atoms, t2))
                    else:
                        docc.append(orb)
                m += 1
    npair = len(sc) // 2
    inp = ValenceInput(0, 0, npair, 0, len(docc), 0, 0, 1 if npair else 0, 0, 0, 0, 0, 0, 0, 0,
                       tol[0], tol[1], tol[2], 0, 0, 0, 0.0, 0.0, [], atom_t, coords,
                       [copy.deepcopy(_O_631G), copy.deepcopy(_H_631G)], [1.0],
                       [[(2 * i + 1, 2 * i + 2) for i in range(npair)]] if npair else [],
                       [], sc + docc)
    return inp.fix_counts()


# LiF rock-salt cluster: the reference's examples/lif128 is truncated
# (114 of 128 geometry lines survive), so it is reconstructed: 4x4x8 lattice, spacing 2.015 Angstrom, x outer /
# z inner, F where ix+iy+iz is even else Li (matches every surviving line, /root/reference/examples/lif128:5-118);
# F basis 6-31+G (numeric content of examples/f-:8-31), Li basis 6-31G (examples/lih.VSHF:8-27); per F five DOCC
# orbitals with the surviving weights (examples/lif128:128-137), per Li one 1s core; tolerances 9 9 8 (lif128:2).
_F_631PG = AtomType(9.0, [
    Shell(0, [7001.71309, 1051.36609, 239.28569, 67.3974453, 21.5199573, 7.4031013],
          [0.0018196169, 0.0139160796, 0.0684053245, 0.23318576, 0.471267439, 0.356618546]),
    Shell(0, [20.8479528, 4.80830834, 1.34406986], [-0.108506975, -0.146451658, 1.12868858]),
    Shell(0, [0.358151393], [1.0]),
    Shell(1, [20.8479528, 4.80830834, 1.34406986], [0.0716287243, 0.345912103, 0.722469957]),
    Shell(1, [0.358151393], [1.0]),
    Shell(0, [0.1076], [1.0]),
    Shell(1, [0.1076], [1.0]),
])
_LI_631G = AtomType(3.0, [
    Shell(0, [642.41892, 96.798515, 22.091121, 6.2010703, 1.9351177, 0.6367358],
          [0.0021426, 0.0162089, 0.0773156, 0.245786, 0.470189, 0.3454708]),
    Shell(0, [2.3249184, 0.6324306, 0.0790534], [-0.0350917, -0.1912328, 1.0839878]),
    Shell(0, [0.035962], [1.0]),
    Shell(1, [2.3249184, 0.6324306, 0.0790534], [0.0089415, 0.1410095, 0.9453637]),
    Shell(1, [0.035962], [1.0]),
])
_F_ORBS = [
    [(1, 1.0)],
    [(2, 0.49077911), (3, 0.55470885), (10, 0.05607011)],
    [(4, 0.62842494), (7, 0.44739668), (11, 0.15779489)],
    [(5, 0.62855015), (8, 0.44737824), (12, 0.15760430)],
    [(6, 0.62862278), (9, 0.44736537), (13, 0.15749744)],
]


def lif_cluster(nx: int = 4, ny: int = 4, nz: int = 8, spacing: float = 2.015,
                tol: Tuple[int, int, int] = (9, 9, 8)) -> ValenceInput:
    """Rock-salt LiF cluster; lif_cluster() is the reconstructed examples/lif128 (128 atoms, 384 DOCC orbitals,
    one-atom orbital basis sets).  Atom order and orbital order follow the surviving part of the file: atoms x
    outer / z inner, orbitals atom by atom."""
    atom_t, coords, docc = [], [], []
    for ix in range(nx):
        for iy in range(ny):
            for iz in range(nz):
                is_f = (ix + iy + iz) % 2 == 0
                atom_t.append(1 if is_f else 2)
                coords.append([ix * spacing, iy * spacing, iz * spacing])
                a = len(atom_t)
                if is_f:
                    for terms in _F_ORBS:
                        docc.append(Orbital([a], list(terms)))
                else:
                    docc.append(Orbital([a], [(1, 1.0)]))
    inp = ValenceInput(0, 0, 0, 0, len(docc), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                       tol[0], tol[1], tol[2], 0, 0, 0, 0.0, 0.0, [], atom_t, coords,
                       [copy.deepcopy(_F_631PG), copy.deepcopy(_LI_631G)], [1.0], [], [], docc)
    return inp.fix_counts()


# n-alkanes from the guess-orbital library of vtools: numeric content of
# /root/reference/vtools/631g/C.basis, H.basis (6-31G) and of the prototype orbitals C_1.gorb (1s core), C-C_1-1.gorb
# (sigma bond from C2H6) and C-H_1-1.gorb (sigma bond from C2H6), which are tabulated for a bond along +z
# (vtools/631g/README-631g): AO index within the two-atom basis set, weight.  vtools rotates them onto every bond of the
# molecule with openbabel's help; here the chain is built directly (all-trans, tetrahedral angles).
_C_631G = AtomType(6.0, [
    Shell(0, [3047.5249, 457.36951, 103.94869, 29.210155, 9.286663, 3.163927],
          [0.0018347, 0.0140373, 0.0688426, 0.2321844, 0.4679413, 0.362312]),
    Shell(0, [7.8682724, 1.8812885, 0.5442493], [-0.1193324, -0.1608542, 1.1434564]),
    Shell(1, [7.8682724, 1.8812885, 0.5442493], [0.0689991, 0.316424, 0.7443083]),
    Shell(0, [0.1687144], [1.0]),
    Shell(1, [0.1687144], [1.0]),
])
_GORB_C_CORE = [(1, 1.0)]
# (AO, weight): carbon AOs 1 s, 2 s, 3-5 p, 6 s, 7-9 p; second atom follows (C: 10-18, H: 10-11); p weights are the z entries
_GORB_CC = {"s": [(2, 0.249), (6, 0.112), (11, 0.249), (15, 0.112)], "pz": [(3, 0.312), (7, 0.168), (12, -0.312), (16, -0.168)]}
_GORB_CH = {"s": [(2, 0.256), (6, 0.130), (10, 0.318), (11, 0.225)], "pz": [(3, 0.336), (7, 0.184)]}


def _bond_orbital(proto: dict, u: np.ndarray) -> List[Tuple[int, float]]:
    """Prototype sigma orbital (bond along +z) rotated onto the unit vector u: s weights unchanged, a p_z weight w becomes
    the p vector w u (first AO of the p shell given in the table)."""
    terms = list(proto["s"])
    for first, w in proto["pz"]:
        for k in range(3):
            if abs(w * u[k]) > 1e-14:
                terms.append((first + k, float(w * u[k])))
    return sorted(terms)


def alkane(n: int, tol: Tuple[int, int, int] = (10, 20, 10), r_cc: float = 1.54, r_ch: float = 1.09) -> ValenceInput:
    """All-trans n-alkane C_n H_(2n+2), 6-31G, one DOCC orbital per core and per bond (VSHF-type wavefunction like
    examples/c3h8): 4n + 1 doubly occupied orbitals with one- and two-atom orbital basis sets."""
    if n < 1:
        raise ValueError("alkane: n >= 1")
    th = np.arccos(-1.0 / 3.0)                         # tetrahedral angle
    sx, cz = np.sin(th / 2.0), np.cos(th / 2.0)
    C = [np.array([i * r_cc * sx, 0.0, (0.5 if i % 2 else -0.5) * r_cc * cz]) for i in range(n)]
    coords: List[np.ndarray] = list(C)
    atom_t = [1] * n
    bonds_ch: List[Tuple[int, int]] = []               # (carbon, hydrogen), 0-based atoms

    def add_h(ic: int, d: np.ndarray):
        coords.append(C[ic] + r_ch * d / np.linalg.norm(d))
        atom_t.append(2)
        bonds_ch.append((ic, len(coords) - 1))

    for i in range(n):
        nb = [C[j] - C[i] for j in (i - 1, i + 1) if 0 <= j < n]
        nb = [v / np.linalg.norm(v) for v in nb]
        if len(nb) == 2:                               # methylene: two H in the plane perpendicular to the backbone
            b = -(nb[0] + nb[1]); b /= np.linalg.norm(b)
            nrm = np.cross(nb[0], nb[1]); nrm /= np.linalg.norm(nrm)
            for sgn in (1.0, -1.0):
                add_h(i, np.cos(th / 2.0) * b + sgn * np.sin(th / 2.0) * nrm)
        else:                                          # methyl (or methane): the remaining tetrahedral directions
            a = nb[0] if nb else np.array([0.0, 0.0, 1.0])
            e0 = np.cross(a, np.array([0.0, 1.0, 0.0]))
            e0 /= np.linalg.norm(e0)
            e1 = np.cross(a, e0)
            for k in range(3):
                ph = 2.0 * np.pi * k / 3.0
                add_h(i, -a / 3.0 + np.sqrt(8.0) / 3.0 * (np.cos(ph) * e0 + np.sin(ph) * e1))
            if not nb:
                add_h(i, a)
    docc: List[Orbital] = []
    for i in range(n):
        docc.append(Orbital([i + 1], list(_GORB_C_CORE)))
    for i in range(n - 1):
        u = (C[i + 1] - C[i]) / np.linalg.norm(C[i + 1] - C[i])
        docc.append(Orbital([i + 1, i + 2], _bond_orbital(_GORB_CC, u)))
    for ic, ih in bonds_ch:
        u = (coords[ih] - C[ic]) / r_ch
        docc.append(Orbital([ic + 1, ih + 1], _bond_orbital(_GORB_CH, u)))
    inp = ValenceInput(0, 0, 0, 0, len(docc), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                       tol[0], tol[1], tol[2], 0, 0, 0, 0.0, 0.0, [], atom_t, [[float(v) for v in x] for x in coords],
                       [copy.deepcopy(_C_631G), copy.deepcopy(_H_631G)], [1.0], [], [], docc)
    return inp.fix_counts()


def dump_json(inp: ValenceInput, path: str, **extra) -> None:
    with open(path, "w") as fh:
        json.dump({"input": inp.to_json(), **extra}, fh, indent=0, separators=(",", ":"))
        fh.write("\n")
