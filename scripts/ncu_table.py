"""Table of key metrics for every kernel launch in an .ncu-rep: python scripts/ncu_table.py rep [out.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = {"gpu__time_duration.sum": "ms", "launch__registers_per_thread": "regs", "launch__occupancy_limit_registers": "occ_reg", "launch__occupancy_limit_shared_mem": "occ_smem",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64%", "smsp__inst_executed.sum": "inst",
        "smsp__sass_inst_executed_op_local_ld.sum": "lld", "smsp__sass_inst_executed_op_local_st.sum": "lst",
        "smsp__sass_inst_executed_op_shared_ld.sum": "sld", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bankc",
        "smsp__sass_inst_executed_op_global_ld.sum": "gld", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "l1tex__t_sector_hit_rate.pct": "l1hit%", "lts__t_sector_hit_rate.pct": "l2hit%"}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
out = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    line = {"kernel": name.replace("void ", "").replace("vb::", "").replace(", ", "|").split("(")[0][:28]}
    for m, short in want.items():
        v = d.get(m, "")
        try:
            v = float(v.replace(",", ""))
            if short == "ms": v = v / (1e6 if rows[1][hdr.index(m)] in ("ns", "nsecond") else 1.0)
            line[short] = f"{v:.3g}"
        except ValueError:
            line[short] = v
    for s in stalls:
        try: line[s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = f"{float(d[s]):.2f}"
        except ValueError: pass
    out.append(line)
keys = list(out[0].keys())
print(",".join(keys))
for l in out: print(",".join(str(l.get(k, "")) for k in keys))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        fh.write(",".join(keys) + "\n")
        for l in out: fh.write(",".join(str(l.get(k, "")) for k in keys) + "\n")
