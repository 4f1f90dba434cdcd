"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Nothing under valence_b200/
does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libvalence_oracle.so")


class Counters(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in (
        "ntasks", "same_orb_skip", "schwarz_pass", "schwarz_erep", "schwarz_exch", "shortcut",
        "int2e_calls", "value_erep", "value_exch", "density2", "shell_quartets",
        "shell_quartets_2e", "determinants", "eri_cached")]

    def asdict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [("enucrep", C.c_double), ("energy", C.c_double), ("wfnorm", C.c_double),
                ("numerator", C.c_double), ("cnt", Counters)]


class FastResult(C.Structure):
    _fields_ = [("enucrep", C.c_double), ("energy", C.c_double), ("wfnorm", C.c_double),
                ("numerator", C.c_double), ("cnt", Counters), ("prim_quartets", C.c_longlong),
                ("blocks", C.c_longlong), ("seconds", C.c_double)]


class RunResult(C.Structure):
    _fields_ = [("enucrep", C.c_double), ("guess_energy", C.c_double), ("total_energy", C.c_double),
                ("converged", C.c_int), ("iterations", C.c_int), ("failed", C.c_int)]


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("valence_oracle.c", "vo_opt.inc", "vo_fast.c", "vo_integrals.c", "vo_internal.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.vo_load.restype = C.c_void_p
        L.vo_load.argtypes = [C.c_char_p]
        L.vo_free.argtypes = [C.c_void_p]
        L.vo_guess_energy.argtypes = [C.c_void_p, C.c_int, C.POINTER(Result)]
        L.vo_first_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(Counters)]
        L.vo_run.argtypes = [C.c_void_p, C.POINTER(RunResult)]
        L.vo_fast_guess_energy.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(FastResult)]
        L.vo_fast_first_order.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(Counters)]
        L.vo_fast_boys.restype = C.c_double
        L.vo_fast_boys.argtypes = [C.c_int, C.c_double]
        L.vo_count_tasks.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.vo_set_quiet.argtypes = [C.c_void_p, C.c_int]
        L.vo_set_memo.argtypes = [C.c_void_p, C.c_int]
        L.vo_set_coords.argtypes = [C.c_void_p, C.c_void_p]
        L.vo_reset_orbitals.argtypes = [C.c_void_p]
        L.vo_guess_partial.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3
        L.vo_baseline_timed.restype = C.c_longlong
        L.vo_baseline_timed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
        L.vo_baseline_sample.restype = C.c_longlong
        L.vo_baseline_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_double)]
        for f in ("vo_nelec", "vo_natom", "vo_norbs", "vo_npairs_schwarz"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.vo_get_wdet.argtypes = [C.c_void_p, C.c_void_p]
        L.vo_get_schwarz.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.vo_get_coeff.argtypes = [C.c_void_p, C.c_void_p]
        L.vo_orbital_eri.restype = C.c_double
        L.vo_orbital_eri.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.vo_orbital_ovl.restype = C.c_double
        L.vo_orbital_ovl.argtypes = [C.c_void_p] + [C.c_int] * 2
        L.vo_orbital_h.restype = C.c_double
        L.vo_orbital_h.argtypes = [C.c_void_p] + [C.c_int] * 2
        _lib = L
    return _lib


class Oracle:
    """One parsed input held by the C oracle."""

    def __init__(self, path: str, quiet: bool = True, memo: bool = True):
        self.L = lib()
        self.h = self.L.vo_load(path.encode())
        if not self.h:
            raise FileNotFoundError(path)
        self.L.vo_set_quiet(self.h, 1 if quiet else 0)
        self.L.vo_set_memo(self.h, 1 if memo else 0)

    def close(self):
        if self.h:
            self.L.vo_free(self.h)
            self.h = None

    __del__ = close

    @property
    def nelec(self):
        return self.L.vo_nelec(self.h)

    @property
    def natom(self):
        return self.L.vo_natom(self.h)

    def guess_energy(self, nrank: int = 1) -> dict:
        r = Result()
        self.L.vo_guess_energy(self.h, nrank, C.byref(r))
        return {"enucrep": r.enucrep, "energy": r.energy, "wfnorm": r.wfnorm,
                "numerator": r.numerator, "counters": r.cnt.asdict()}

    def fast_guess_energy(self, nthreads: int = 0, thr: float = 0.0) -> dict:
        """guess_energy through the fast (inverse-form, AO-block) oracle path, vo_fast.c."""
        r = FastResult()
        rc = self.L.vo_fast_guess_energy(self.h, nthreads, thr, C.byref(r))
        if rc != 0:
            raise RuntimeError("fast oracle: input outside its limits (singular overlap block, more than 12 spin-coupled pairs)")
        return {"enucrep": r.enucrep, "energy": r.energy, "wfnorm": r.wfnorm, "numerator": r.numerator,
                "counters": r.cnt.asdict(), "prim_quartets": int(r.prim_quartets), "seconds": r.seconds}

    def fast_first_order(self, iorb: int, nthreads: int = 0, thr: float = 0.0):
        """ham, ovl of first_order_opt for 1-based orbital iorb through the fast path."""
        hd = C.c_int(0)
        cnt = Counters()
        big = 64
        ham = np.zeros(big * big)
        ovl = np.zeros(big * big)
        n = self.L.vo_fast_first_order(self.h, iorb, nthreads, thr, ham.ctypes.data, ovl.ctypes.data, C.byref(hd), C.byref(cnt))
        if n < 0:
            raise RuntimeError("fast oracle: input outside its limits")
        h = hd.value
        return ham[:h * h].reshape(h, h).T[:n, :n].copy(), ovl[:h * h].reshape(h, h).T[:n, :n].copy()

    def count_tasks(self) -> dict:
        """Screening / task counters of guess_energy without the 2e integrals and determinants."""
        cnt = Counters()
        self.L.vo_count_tasks(self.h, C.byref(cnt))
        return cnt.asdict()

    def wdet(self) -> np.ndarray:
        n = self.nelec
        a = np.zeros((n, n))
        self.L.vo_get_wdet(self.h, a.ctypes.data)
        return a.T.copy()  # column-major in C

    def schwarz(self) -> np.ndarray:
        n = self.L.vo_npairs_schwarz(self.h)
        a = np.zeros(n)
        self.L.vo_get_schwarz(self.h, a.ctypes.data, n)
        return a

    def first_order(self, iorb: int):
        """ham, ovl (norbas x norbas) of first_order_opt for 1-based orbital iorb."""
        hd = C.c_int(0)
        cnt = Counters()
        big = 64
        ham = np.zeros(big * big)
        ovl = np.zeros(big * big)
        n = self.L.vo_first_order(self.h, iorb, ham.ctypes.data, ovl.ctypes.data, C.byref(hd), C.byref(cnt))
        h = hd.value
        H = ham[:h * h].reshape(h, h).T[:n, :n].copy()
        S = ovl[:h * h].reshape(h, h).T[:n, :n].copy()
        return H, S, cnt.asdict()

    def run(self) -> dict:
        r = RunResult()
        rc = self.L.vo_run(self.h, C.byref(r))
        return {"rc": rc, "enucrep": r.enucrep, "guess_energy": r.guess_energy,
                "total_energy": r.total_energy, "converged": bool(r.converged),
                "iterations": r.iterations, "failed": r.failed}

    def set_coords(self, x_angstrom):
        x = np.ascontiguousarray(x_angstrom, dtype=np.float64).ravel()
        self.L.vo_set_coords(self.h, x.ctypes.data)

    def guess_partial(self, irank: int, nrank: int):
        e, w, n = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
        self.L.vo_guess_partial(self.h, irank, nrank, C.byref(e), C.byref(w), C.byref(n))
        return e.value, w.value, n.value

    def baseline_sample(self, irank: int, nrank: int, task_limit: int):
        e = C.c_double(0.0)
        nq = self.L.vo_baseline_sample(self.h, irank, nrank, task_limit, C.byref(e))
        return int(nq), e.value

    def baseline_timed(self, irank: int, nrank: int, seconds: float):
        el = C.c_double(0.0)
        nq = self.L.vo_baseline_timed(self.h, irank, nrank, seconds, C.byref(el))
        return int(nq), el.value


def _baseline_worker(args):
    path, irank, nrank, seconds = args
    o = Oracle(path, memo=False)
    nq, el = o.baseline_timed(irank, nrank, seconds)
    o.close()
    return nq, el


def cpu_baseline(path: str, seconds: float = 12.0, nproc: int = 0) -> dict:
    """The reference algorithm (literal restatement, no memoisation) on the host cores with the
    reference's own decomposition: rank r of nproc takes tasks r, r+nproc, ... (valence.F90:1162-1163).
    Every worker runs for `seconds`; throughput = shell quartets / max elapsed."""
    import multiprocessing as mp
    nproc = nproc or (os.cpu_count() or 1)
    build()
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        res = pool.map(_baseline_worker, [(path, r, nproc, seconds) for r in range(nproc)])
    nq = sum(r[0] for r in res)
    el = max(r[1] for r in res)
    return {"shell_quartets": nq, "seconds": el, "quartets_per_s": nq / el if el > 0 else 0.0, "cores": nproc}
