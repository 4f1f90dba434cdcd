#!/usr/bin/env python
"""Benchmark of the VSVB energy hot path (BASELINE.json metric: contracted shell quartets / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--sweep 16,32,64,128] [--impl ours|reference]

One step = one pass of the hot path through the library's public call: geometry in from (pinned) host memory,
guess_energy (one vsvb_energy evaluation: AO one-electron matrices, spin-block inverses, pair tables, Schwarz pass,
fused ERI + contraction pass), energy out to the host.  Workloads:
    w256 (default)  synthetic (H2O)_256, 6-31G, five DOCC orbitals per monomer -- the cluster BASELINE.json's north_star
                    names (768 atoms, 3328 AOs, 1280 doubly occupied orbitals); tolerances `10 20 10`
    w<N>            the same generator at N molecules
    lif128          the reconstructed examples/lif128 (4x4x8 LiF lattice, 384 DOCC orbitals), tolerances 9 20 8
    h2o, c3h8, cu+.3d94s1   the reference's example inputs verbatim (configs 1-3 of BASELINE.json)
The unit counted is the reference's: one contracted AO shell quartet evaluated and digested = one simint_compute_eri
call of the reference algorithm (/root/reference/src/valence.F90:3398); the count comes from the engine's bit-exact
screening counters, so recomputation the GPU formulation avoids still counts once per reference call.
`primitive_quartets_per_step` is what the GPU actually evaluates.

`parity` compares the step's energy and counters with the fast CPU oracle's committed result for the same input
(tests/golden/fast__*.json) when there is one.

Metric 2 (`energy_plus_first_order`): wall time of one guess energy plus the first_order_opt matrices (ham, ovl) of
orbital 1 on the same input.

N > 1 (torchrun): one process per GPU; the ranks form an NCCL communicator INSIDE the library (vb_nccl.cpp), the tile
list is split by bra blocks, the table shares are all-gathered over NVLink and ONE all-reduce sums the packed
accumulators per step; "scaling": "strong" (the same input is split over more GPUs).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contracted_shell_quartets_per_s"
UNIT = "shell quartets/s"
EXAMPLES = {"h2o": "examples__h2o", "c3h8": "examples__c3h8", "cu+.3d94s1": "examples__cu+.3d94s1"}


def workload_input(name: str):
    """(ValenceInput, description, fixture case or None)"""
    from valence_b200 import inputs
    if name.startswith("w") and name[1:].isdigit():
        n = int(name[1:])
        return inputs.water_cluster(n, tol=(10, 20, 10)), f"(H2O)_{n} 6-31G VSVB guess energy, tolerances 10 20 10", name
    if name == "lif128":
        return inputs.lif_cluster(tol=(9, 20, 8)), "LiF 4x4x8 lattice (reconstructed examples/lif128) VSVB guess energy, tolerances 9 20 8", "lif128"
    if name in EXAMPLES:
        with open(os.path.join(ROOT, "tests", "golden", EXAMPLES[name] + ".json")) as fh:
            d = json.load(fh)
        return inputs.ValenceInput.from_json(d["input"]), f"examples/{name} (reference input verbatim) VSVB guess energy", None
    raise SystemExit(f"unknown workload {name}")


def make_input(name: str):
    from valence_b200 import inputs
    inp, desc, case = workload_input(name)
    fd, path = tempfile.mkstemp(prefix=f"vb_{name.replace('/', '_')}_", suffix=".inp")
    with os.fdopen(fd, "w") as fh:
        fh.write(inputs.write(inp))
    return path, inp, desc, case


def fixture(case):
    if not case:
        return None
    try:
        with open(os.path.join(ROOT, "tests", "golden", f"fast__{case}.json")) as fh:
            return json.load(fh)
    except OSError:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> dict:
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def traffic_from_profiles(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the class-pass launches for this workload, summed over the launches of
    one energy pass, from the committed ncu capture (profiles/r2_pclass_dram_<workload>.csv); None when there is none."""
    path = os.path.join(ROOT, "profiles", f"r2_pclass_dram_{workload}.csv")
    try:
        tot = 0.0
        with open(path) as fh:
            for line in fh:
                f = [x.strip().strip('"') for x in line.split(",")]
                if len(f) >= 3 and f[1].startswith("dram__bytes_") and f[1].endswith(".sum"):
                    tot += float(f[2])
        return tot or None
    except OSError:
        return None


def cpu_sample_label(seconds: float, cores: int) -> str:
    return (f"{seconds:.0f} s of the reference's own task list on the same input -- this rank's share of schwarz_ints "
            f"(valence.F90:1489-1523) and then of the 2e loop (:1163-1433), literal C restatement, no AO memoisation, no integral "
            f"cache, round-robin over {cores} processes as the reference's MPI ranks; for the large clusters the sample ends inside "
            f"schwarz_ints, i.e. it times int2e shell-quartet loops only and no n^3 Givens determinant of the 2e loop")


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU algorithm (literal restatement under oracle/,
    since the Fortran/SIMINT reference cannot be built here) on all host cores, on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    path, inp, desc, case = make_input(args.workload)
    cores = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    per_step = float(os.environ.get("VB_BENCH_REF_SECONDS", per_step))      # tests shorten the sample
    vals = []
    for i in range(args.warmup + args.steps):
        r = oracle.cpu_baseline(path, seconds=per_step, nproc=cores)
        if i >= args.warmup:
            vals.append(r)
    nq = sum(v["shell_quartets"] for v in vals)
    sec = sum(v["seconds"] for v in vals)
    value = nq / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "name": args.workload},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_label(per_step, cores) + " (per step)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    os.unlink(path)


def timed_steps(eng, x_host, steps, warmup, barrier):
    for _ in range(warmup):
        eng.set_coords(x_host.numpy())
        eng.energy()
    barrier()
    t0 = time.perf_counter()
    res = []
    for _ in range(steps):
        eng.set_coords(x_host.numpy())
        res.append(eng.energy())
    barrier()
    return res, time.perf_counter() - t0


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from valence_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))     # only for the timing barriers / max over ranks
        # all ranks of the node build their share of the host tables at the same time: share the cores
        os.environ.setdefault("VB_HOST_THREADS", str(max(2, (os.cpu_count() or 2) // world)))
    torch.cuda.set_device(local)

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    def red(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    R = dist.ReduceOp if world > 1 else None
    MAX, SUM = (R.MAX, R.SUM) if R else (None, None)
    # job-unique part of the library's NCCL rendezvous keys: rank 0's pid and start time, agreed through the all-reduce
    comm_seq = [0]
    job_token = "%d_%d" % (int(red(float(os.getpid()) if rank == 0 else 0.0, SUM)), int(red(float(int(time.time())) if rank == 0 else 0.0, SUM)))
    peak = api.measure_fp64_peak(local)

    def measure(name, steps, warmup, sample_clocks):
        """Time `steps` steps of workload `name`; returns the per-workload record (rank 0) and the engine."""
        path, inp, desc, case = make_input(name)
        eng = api.Engine(path, device=local)
        if world > 1:
            # NCCL inside the library: energy() is collective from here on.  One rendezvous key per engine and job.
            comm_seq[0] += 1
            eng.attach_comm(rank, world, f"bench{job_token}_{comm_seq[0]}")
        x_host = torch.tensor(inp.coords, dtype=torch.float64).flatten().pin_memory()
        sampler = ClockSampler(local) if (rank == 0 and sample_clocks) else None
        if sampler:
            # warm-up happens inside timed_steps before its barrier; the sampler covers warm-up + timed region
            sampler.start()
        results, wall = timed_steps(eng, x_host, steps, warmup, barrier)
        if sampler:
            sampler.stop_flag.set()
            sampler.join(timeout=3)
        wall = red(wall, MAX)
        dev_ms = red(sum(r["t_1e_ms"] + r["t_density_ms"] + r["t_diag_ms"] + r["t_tiles_ms"] for r in results) / steps, MAX)
        tile_ms = red(sum(r["t_tiles_ms"] for r in results) / steps, MAX)
        flops_rank = sum(r["flops_model"] for r in results) / steps
        tflops_rank = sum(r.get("flops_transform", 0.0) for r in results) / steps
        tile_ms_rank = sum(r["t_tiles_ms"] for r in results) / steps
        launches = red(float(sum(r["launches"] for r in results)), SUM)
        primq = red(sum(r["n_prim_quartets"] for r in results) / steps, SUM)
        farq = red(sum(r["n_ao_quartets"] for r in results) / steps, SUM)
        last = results[-1]
        refq = last["ref_shell_quartets"]                # all-reduced inside the library: the whole job's count
        e2e_ms = 1e3 * wall / steps
        achieved = flops_rank / (tile_ms_rank * 1e-3) / 1e12 if tile_ms_rank > 0 else 0.0
        achieved_t = tflops_rank / (tile_ms_rank * 1e-3) / 1e12 if tile_ms_rank > 0 else 0.0
        rec = {"name": name, "workload": desc, "natom": eng.natom, "electrons": eng.nelec, "ms_per_step": dev_ms, "wall_ms_per_step": e2e_ms,
               "tile_pass_ms_per_step": tile_ms, "value": refq / (dev_ms * 1e-3), "e2e_value": refq / (e2e_ms * 1e-3),
               "energy_hartree": last["energy"], "reference_algorithm_shell_quartets_per_step": refq,
               "primitive_quartets_per_step": primq, "far_field_primitive_quartets_per_step": farq,
               "h2d_bytes_per_step": int(last.get("h2d_bytes", 0)) or 24 * eng.natom, "d2h_bytes_per_step": int(last.get("d2h_bytes", 0)) or 8,
               "gpu_launches": int(launches), "roofline_achieved_tflops": achieved, "roofline_frac": achieved / peak if peak else None,
               "transform_tflops": achieved_t,
               "clocks": sampler.summary() if sampler else None}
        fx = fixture(case)
        if fx is not None:
            cnt = last["counters"]
            keys = ("schwarz_erep", "schwarz_exch", "int2e_calls", "shell_quartets_2e", "shortcut", "value_erep", "value_exch")
            rec["parity"] = {"oracle": "oracle/vo_fast.c, committed as tests/golden/fast__%s.json" % case, "oracle_energy": fx["energy"],
                             "dE_hartree": last["energy"] - fx["energy"], "counters_identical": all(cnt[k] == fx["counters"][k] for k in keys),
                             "oracle_cpu_seconds": fx.get("seconds"), "oracle_threads": fx.get("threads")}
        return rec, eng, path, inp

    rec, eng, path, inp = measure(args.workload, args.steps, args.warmup, True)

    # Metric 2: guess energy + one first_order_opt accumulator build (ham, ovl of orbital 1) on the same input.
    grad = None
    os.environ.setdefault("VB_FO_REQUIRE_CACHE", "1")      # never fall back to norbas^2/2 full tile passes inside the bench
    if not args.no_grad:
        barrier()
        t0g = time.perf_counter()
        try:
            rg = eng.energy()
            Hg, Sg, st = eng.first_order(1)
            barrier()
            tg = red(time.perf_counter() - t0g, MAX)
            import numpy as np
            cw = np.array([w for _, w in inp.orbitals[0].terms])
            grad = {"ms": 1e3 * tg, "orbital": 1, "matrix_order": int(Hg.shape[0]),
                    "first_order_kernel_ms": red(st["t_tiles_ms"], MAX), "kernel_launches": int(st["launches"]),
                    "method": "rank-one form: one Fock-like matrix pass over the tiles free of the subject orbital + the subject tiles per element",
                    "rayleigh_quotient_minus_energy": float(cw @ Hg @ cw / (cw @ Sg @ cw)) + rg["enucrep"] - rg["energy"]}
        except RuntimeError as ex:      # metric 2 must never cost the headline line
            grad = {"error": str(ex)[:200]}
    eng.close()

    # scaling sweep over cluster sizes (config 5): short runs, same call path
    sweep = []
    for n in args.sweep:
        try:
            r2, e2, p2, _ = measure(f"w{n}", 2, 1, False)
        except Exception as ex:
            if world > 1:               # collective: the other ranks are inside the same call
                raise
            sweep.append({"name": f"w{n}", "error": str(ex)[:200]})      # a side block must never cost the headline line
            continue
        e2.close()
        os.unlink(p2)
        if rank == 0:
            sweep.append({k: r2[k] for k in ("name", "ms_per_step", "wall_ms_per_step", "tile_pass_ms_per_step", "value", "e2e_value", "energy_hartree",
                                             "primitive_quartets_per_step", "roofline_achieved_tflops", "roofline_frac") if k in r2} | ({"parity": r2["parity"]} if "parity" in r2 else {}))

    # the other stated configurations (BASELINE.json configs 1-4) through the same call path: short runs, one GPU
    configs = []
    if world == 1 and args.sweep:
        for cname in ("h2o", "c3h8", "cu+.3d94s1", "lif128"):
            try:
                rc, ec, pc, _ = measure(cname, 2, 1, False)
                ec.close()
                ent = {k: rc[k] for k in ("name", "workload", "ms_per_step", "wall_ms_per_step", "tile_pass_ms_per_step", "value", "e2e_value", "energy_hartree",
                                          "reference_algorithm_shell_quartets_per_step", "primitive_quartets_per_step", "roofline_achieved_tflops",
                                          "roofline_frac") if k in rc}
                if "parity" in rc:
                    ent["parity"] = rc["parity"]
                if cname in EXAMPLES:
                    with open(os.path.join(ROOT, "tests", "golden", EXAMPLES[cname] + ".json")) as fh:
                        ent["reference_golden_guess_energy"] = json.load(fh)["golden"]["guess_energy"]
                if cname in ("h2o", "cu+.3d94s1"):       # the literal restatement of the reference runs these to completion in seconds
                    from oracle.oracle import Oracle
                    o = Oracle(pc)
                    t0 = time.perf_counter()
                    ro = o.guess_energy()
                    dtc = time.perf_counter() - t0
                    o.close()
                    ent["cpu_literal_oracle"] = {"seconds": dtc, "cores": 1, "energy_hartree": ro["energy"],
                                                 "shell_quartets_per_s": ro["counters"]["shell_quartets_2e"] / dtc if dtc > 0 else None}
                configs.append(ent)
                os.unlink(pc)
            except Exception as ex:      # a side block must never cost the headline line
                configs.append({"name": cname, "error": str(ex)[:200]})

    # config 5's spin-coupled variant: the two OH bonds of the first two molecules of (H2O)_32 as spin-coupled pairs (one GPU)
    scv = None
    if world == 1 and args.sweep:
        from valence_b200 import inputs
        fd, psc = tempfile.mkstemp(prefix="vb_w32sc2_", suffix=".inp")
        with os.fdopen(fd, "w") as fh:
            fh.write(inputs.write(inputs.water_cluster(32, tol=(10, 20, 10), sc_molecules=2)))
        try:
            esc = api.Engine(psc, device=local)
            esc.energy()
            t0 = time.perf_counter()
            rs = esc.energy()
            scv = {"workload": "(H2O)_32 6-31G, OH bonds of the first 2 molecules as spin-coupled pairs: npair 4, 256 determinant pairs, "
                               "spin blocks of 160 inverted on the GPU per pair", "wall_ms": 1e3 * (time.perf_counter() - t0),
                   "energy_hartree": rs["energy"], "kernel_launches": int(rs["launches"])}
            esc.close()
        except RuntimeError as ex:
            scv = {"error": str(ex)[:200]}
        os.unlink(psc)

    if rank == 0:
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": rec["workload"], "name": args.workload, "atoms": rec["natom"], "electrons": rec["electrons"],
                       "l2": "every step rebuilds and re-uploads its tables (0.6 GB for (H2O)_256, > L2) and streams 38 GB of "
                             "orbital-level integral blocks through HBM/L2: no warm-cache advantage between steps",
                       "parallelism": f"{world} GPU(s): tile list split by bra blocks, table shares all-gathered over NVLink, 1 NCCL all-reduce/step (in-library)"},
            "energy_hartree": rec["energy_hartree"],
            "reference_algorithm_shell_quartets_per_step": rec["reference_algorithm_shell_quartets_per_step"],
            "primitive_quartets_per_step": rec["primitive_quartets_per_step"],
            "far_field_primitive_quartets_per_step": rec["far_field_primitive_quartets_per_step"],
            "wall_ms_per_step": rec["wall_ms_per_step"],
            "tile_pass_ms_per_step": rec["tile_pass_ms_per_step"],
            "e2e": {"value": rec["e2e_value"], "unit": UNIT, "ms_per_step": rec["wall_ms_per_step"],
                    "h2d_bytes_per_step": rec["h2d_bytes_per_step"], "d2h_bytes_per_step": rec["d2h_bytes_per_step"],
                    "note": "valence_api-style call: geometry in from (pinned) host memory, energy out to host; all basis / pair tables "
                            "are rebuilt on the host and copied to the device inside the timed region"},
            "gpu_launches": rec["gpu_launches"],
            "roofline": {"bound": "fp64", "achieved": rec["roofline_achieved_tflops"], "peak": peak, "unit": "TFLOP/s", "frac": rec["roofline_frac"],
                         "traffic": traffic_from_profiles(args.workload) if world == 1 else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the class + contraction launches of one energy pass, ncu capture of "
                                           "the same workload on ONE GPU committed as profiles/r2_pclass_dram_<workload>.csv (not measured in this run; null "
                                           "when there is none and on several GPUs, where a rank's launches move its share only)",
                         "transforms": {"achieved": rec["transform_tflops"], "unit": "TFLOP/s",
                                        "frac_integrals_plus_transforms": (rec["roofline_achieved_tflops"] + rec["transform_tflops"]) / peak if peak else None,
                                        "note": "executed FP64 tensor-core flops (DMMA m8n8k4 = 512 flops, in-kernel instruction count) of the two density "
                                                "transforms that contract every integral block with the orbital-pair densities inside the same kernels; "
                                                "DMMA and DFMA share one FP64 pipe on B200 (37 vs 36 TFLOP/s measured), so the second fraction is the pipe's "
                                                "useful load.  `achieved` / `frac` above count the integral arithmetic only"},
                         "kernel": "tile pass = k_pclass<TB,TK> x 9 integral classes (fused primitive ERI + DMMA density transforms, FP64 reductions "
                                   "of the orbital-level blocks at L2) + k_contract_items (screening + cofactor contraction)",
                         "peak_source": "measured here: DFMA micro-benchmark vb_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)",
                         "flops": "algorithmic, as executed: primitive quartets per class from in-kernel counters x the class's operation count, "
                                  "quartets in the asymptotic regime (T >= 40) at their shorter count; the DMMA transform flops are not counted (DESIGN.md)"},
            "clocks": rec["clocks"],
        }
        if "parity" in rec:
            line["parity"] = rec["parity"]
        if grad is not None:
            line["energy_plus_first_order"] = grad
        if sweep:
            line["sweep"] = sweep
        if scv is not None:
            line["spin_coupled_variant"] = scv
        if configs:
            line["configs"] = configs
        if args.cpu_baseline_seconds > 0 and world == 1:
            from oracle import oracle
            cb = oracle.cpu_baseline(path, seconds=args.cpu_baseline_seconds)
            line["cpu_baseline"] = {"value": cb["quartets_per_s"], "unit": UNIT, "cores": cb["cores"], "kind": "port",
                                    "sample": cpu_sample_label(cb["seconds"], cb["cores"])}
        print(json.dumps(line), flush=True)
    os.unlink(path)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("VB_BENCH_WORKLOAD", "w256"))
    ap.add_argument("--waters", type=int, default=0, help="shorthand for --workload w<N>")
    ap.add_argument("--sweep", default=os.environ.get("VB_BENCH_SWEEP"),
                    help="cluster sizes of the scaling sweep reported in `sweep` ('' = none); default 16,32,64,128 on one GPU, none on "
                         "several (the sweep and the other stated configurations are single-GPU lines; the N-GPU line is the headline workload)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-grad", action="store_true", help="skip metric 2 (energy + first_order_opt matrices)")
    ap.add_argument("--grad-waters", type=int, default=-1, help="(kept for compatibility) 0 = skip metric 2")
    args = ap.parse_args()
    if args.waters:
        args.workload = f"w{args.waters}"
    if args.grad_waters == 0:
        args.no_grad = True
    if args.sweep is None:
        args.sweep = "16,32,64,128" if int(os.environ.get("WORLD_SIZE", "1")) == 1 else ""
    args.sweep = [int(x) for x in str(args.sweep).split(",") if x.strip()]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
