#!/usr/bin/env python
"""Executed scalar FP64 operations per kernel from an `ncu --set full` report.

    python scripts/ncu_fp64_exec.py gpurun_out/r2b_pclass_n64.ncu-rep > profiles/r2b_pclass_fp64_executed_H2O64.csv

The report holds the thread-level DADD / DMUL / DFMA counts as rates per elapsed cycle; rate x elapsed cycles is the count.
DMMA (the density transforms) is not in these counters: the engine counts it itself (PQ_DMMA in vb_pclass.cuh).
"""
import csv
import subprocess
import sys

OPS = ("dadd", "dmul", "dfma")


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    col = {o: h.index("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o) for o in OPS}
    name, dur, cyc = h.index("Kernel Name"), h.index("gpu__time_duration.sum"), h.index("smsp__cycles_elapsed.avg")
    print("kernel,ms,dadd,dmul,dfma,fp64_flops")
    tot = 0.0
    for r in rows[2:]:
        c = float(r[cyc])
        n = {o: float(r[col[o]]) * c for o in OPS}
        fl = n["dadd"] + n["dmul"] + 2 * n["dfma"]
        tot += fl
        k = r[name].replace("void ", "").replace("(TileArgs, ClassCfg)", "").replace(", ", "|")
        print("%s,%.2f,%.4g,%.4g,%.4g,%.4g" % (k, float(r[dur]), n["dadd"], n["dmul"], n["dfma"], fl))
    print("total,,,,,%.4g" % tot)


if __name__ == "__main__":
    main(sys.argv[1])
