"""Python host-side binding of libvalence_b200.so (ctypes over the C-ABI in
include/valence_b200.h).  Mirrors the reference's library interface
(/root/reference/src/valence_api.F90:9,37,114): initialise from an input
file, calculate the energy for a geometry, finalise.

The CUDA library is mandatory: importing succeeds without it, but creating an
Engine raises when the library cannot be built/loaded or no GPU is visible.
There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import build as _build
from . import distributed as _dist

CNT_NAMES = ("schwarz_erep", "schwarz_exch", "value_erep", "value_exch", "int2e_calls",
             "shell_quartets_2e", "shortcut", "entries")


class CEnergyResult(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ("energy", "enucrep", "numerator", "wfnorm", "e1", "e2")]
                + [("counters", C.c_longlong * len(CNT_NAMES))]
                + [(n, C.c_longlong) for n in ("n_entries", "n_groups", "n_pairgroups", "n_tiles",
                                               "n_tiles_mine", "n_ao_quartets", "n_prim_quartets")]
                + [("flops_model", C.c_double), ("ref_shell_quartets", C.c_longlong)]
                + [(n, C.c_double) for n in ("t_total_ms", "t_host_setup_ms", "t_1e_ms", "t_density_ms",
                                             "t_diag_ms", "t_tiles_ms")]
                + [(n, C.c_int) for n in ("launches", "diag_launches", "tile_launches")]
                + [("min_pivot_ratio", C.c_double), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("flops_transform", C.c_double)])

    def asdict(self) -> dict:
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            if name == "counters":
                d[name] = {k: int(v[i]) for i, k in enumerate(CNT_NAMES)}
            else:
                d[name] = v
        return d


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_needed: bool = True):
    """Load the CUDA library; raise loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_needed and not os.environ.get("VB_LIB_SUFFIX") and _build.stale():   # experiment variants are prebuilt
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError("libvalence_b200.so is missing: the CUDA extension is required (no CPU path)")
    L = C.CDLL(_build.LIB, mode=C.RTLD_GLOBAL)
    L.vb_last_error.restype = C.c_char_p
    L.vb_engine_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.vb_engine_destroy.argtypes = [C.c_void_p]
    L.vb_engine_natom.argtypes = [C.c_void_p]
    L.vb_engine_nelec.argtypes = [C.c_void_p]
    L.vb_engine_norbas.argtypes = [C.c_void_p, C.c_int]
    L.vb_engine_attach_comm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    L.vb_engine_attach_nccl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.vb_engine_set_coords.argtypes = [C.c_void_p, C.c_void_p]
    L.vb_engine_energy.argtypes = [C.c_void_p, C.POINTER(CEnergyResult)]
    L.vb_engine_energy_partial.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(CEnergyResult)]
    L.vb_engine_energy_finish.argtypes = [C.c_void_p, C.POINTER(CEnergyResult)]
    L.vb_engine_accum_device.restype = C.c_void_p
    L.vb_engine_accum_device.argtypes = [C.c_void_p]
    L.vb_engine_accum_len.argtypes = [C.c_void_p]
    L.vb_engine_stream.restype = C.c_void_p
    L.vb_engine_stream.argtypes = [C.c_void_p]
    L.vb_engine_first_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(CEnergyResult)]
    L.vb_engine_shard_tables.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    L.vb_engine_first_order_sharded.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                                C.POINTER(CEnergyResult)]
    L.vb_engine_run.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int)] * 2
    L.vb_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


class _DeviceView:
    """Expose the engine's accumulator as a CUDA array for torch.as_tensor()."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


class Engine:
    """One VSVB problem bound to one GPU (valence_api_initialize)."""

    def __init__(self, input_path: str, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.vb_engine_create(os.fspath(input_path).encode(), device, C.byref(h))
        if rc != 0:
            raise RuntimeError(self.L.vb_last_error().decode())
        self.h = h
        self.device = device
        self.comm = None      # (rank, nranks) once the library holds an NCCL communicator (attach_comm)

    def close(self):
        if getattr(self, "h", None):
            self.L.vb_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.vb_last_error().decode())

    @property
    def natom(self) -> int:
        return self.L.vb_engine_natom(self.h)

    @property
    def nelec(self) -> int:
        return self.L.vb_engine_nelec(self.h)

    def norbas(self, iorb: int) -> int:
        """Number of expansion terms of 1-based orbital iorb (order of its first_order_opt matrices)."""
        n = self.L.vb_engine_norbas(self.h, iorb)
        if n < 0:
            raise RuntimeError(f"no orbital {iorb}")
        return n

    def attach_comm(self, rank: int, nranks: int, key: Optional[str] = None):
        """One process per GPU: the ranks of the job form an NCCL communicator INSIDE the library (no torch.distributed
        needed); energy(), first_order() and run() are collective afterwards.  Replaces xm_propagate / xm_equalize*."""
        self._check(self.L.vb_engine_attach_comm(self.h, rank, nranks, (key or "").encode()))
        self.comm = (rank, nranks) if nranks > 1 else None

    def set_coords(self, x_angstrom: Sequence[float]):
        x = np.ascontiguousarray(x_angstrom, dtype=np.float64).ravel()
        assert x.size == 3 * self.natom
        self._check(self.L.vb_engine_set_coords(self.h, x.ctypes.data))

    def energy(self, x_angstrom: Optional[Sequence[float]] = None) -> dict:
        """valence_api_calculate_energy for one GPU: host buffers in, host numbers out."""
        if x_angstrom is not None:
            self.set_coords(x_angstrom)
        r = CEnergyResult()
        self._check(self.L.vb_engine_energy(self.h, C.byref(r)))
        return r.asdict()

    def first_order(self, iorb: int):
        """ham, ovl (norbas x norbas) of first_order_opt for 1-based orbital iorb, plus run statistics."""
        cap = self.norbas(iorb) ** 2
        ham = np.zeros(cap)
        ovl = np.zeros(cap)
        n = C.c_int(0)
        r = CEnergyResult()
        self._check(self.L.vb_engine_first_order(self.h, iorb, ham.ctypes.data, ovl.ctypes.data, cap, C.byref(n), C.byref(r)))
        k = n.value
        return ham[:k * k].reshape(k, k).T.copy(), ovl[:k * k].reshape(k, k).T.copy(), r.asdict()

    def first_order_distributed(self, iorb: int, rank: int, nranks: int):
        """first_order_opt matrices with the tiles sharded over the ranks and ONE all-reduce of ham
        (the reference equalizes ham the same way, valence.F90:755-764)."""
        if self.comm is not None:          # the library's own communicator shards and sums
            return self.first_order(iorb)
        cap = self.norbas(iorb) ** 2
        ham = np.zeros(cap)
        ovl = np.zeros(cap)
        n = C.c_int(0)
        r = CEnergyResult()
        err = None
        try:
            self._check(self.L.vb_engine_first_order_sharded(self.h, iorb, rank, nranks, ham.ctypes.data, ovl.ctypes.data, cap, C.byref(n),
                                                             C.byref(r)))
        except RuntimeError as ex:      # keep the ranks in step: agree on failure before the all-reduce of ham
            err = ex
        k = n.value
        if nranks > 1:
            import torch
            import torch.distributed as dist
            ok = torch.tensor([0.0 if err else 1.0], dtype=torch.float64, device=f"cuda:{self.device}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 0.5:
                raise RuntimeError(f"first_order failed on a rank: {err}" if err else "first_order failed on another rank")
            t = torch.from_numpy(ham[:k * k].copy()).to(f"cuda:{self.device}")
            _dist.allreduce_sum(t)
            ham[:k * k] = t.cpu().numpy()
        elif err:
            raise err
        return ham[:k * k].reshape(k, k).T.copy(), ovl[:k * k].reshape(k, k).T.copy(), r.asdict()

    def run(self, print_output: bool = False) -> dict:
        """calculate_vsvb_energy: guess energy + (if requested by the input) orbital / spin optimisation."""
        nuc, g, t = C.c_double(0), C.c_double(0), C.c_double(0)
        conv, it = C.c_int(0), C.c_int(0)
        self._check(self.L.vb_engine_run(self.h, 1 if print_output else 0, C.byref(nuc), C.byref(g), C.byref(t), C.byref(conv), C.byref(it)))
        return {"enucrep": nuc.value, "guess_energy": g.value, "total_energy": t.value, "converged": bool(conv.value), "iterations": it.value}

    # --- sharded form (one process per GPU) ---------------------------------
    def energy_partial(self, rank: int, nranks: int) -> CEnergyResult:
        r = CEnergyResult()
        self._check(self.L.vb_engine_energy_partial(self.h, rank, nranks, C.byref(r)))
        return r

    def accumulator(self):
        """torch CUDA tensor aliasing the packed accumulators [E2, counters...]."""
        import torch
        n = self.L.vb_engine_accum_len(self.h)
        ptr = self.L.vb_engine_accum_device(self.h)
        return torch.as_tensor(_DeviceView(ptr, n), device=f"cuda:{self.device}")

    def energy_finish(self, r: CEnergyResult) -> dict:
        self._check(self.L.vb_engine_energy_finish(self.h, C.byref(r)))
        return r.asdict()

    def energy_distributed(self, rank: int, nranks: int) -> dict:
        """Shard the tile list over ranks, one NCCL all-reduce of the accumulators.

        VB_SHARD_SETUP=1 (ranks of ONE node): the host-side pair tables are built once per node instead of once per
        rank -- every rank publishes its share under /dev/shm, a barrier, then each rank merges all shares."""
        if self.comm is not None:          # the library's own communicator: one collective call
            return self.energy()
        if nranks > 1 and os.environ.get("VB_SHARD_SETUP", "1") != "0":
            import torch.distributed as dist
            key = os.environ.get("VB_SHARD_KEY") or f"{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}"
            prefix = f"/dev/shm/valence_b200_{key}_"
            self._check(self.L.vb_engine_shard_tables(self.h, rank, nranks, prefix.encode()))
            dist.barrier()
        r = self.energy_partial(rank, nranks)
        if nranks > 1:
            import torch
            import torch.distributed as dist
            acc = self.accumulator()
            torch.cuda.synchronize(self.device)
            _dist.allreduce_sum(acc)
            torch.cuda.synchronize(self.device)
        return self.energy_finish(r)


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 FMA peak (TFLOP/s) of one GPU: the ERI kernel's roofline denominator."""
    L = load()
    v = C.c_double(0.0)
    if L.vb_measure_fp64_peak(device, C.byref(v)) != 0:
        raise RuntimeError(L.vb_last_error().decode())
    return v.value
