#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel from an `ncu --set full --import-source on` report.

    ncu -i <report> --page source --csv --print-source cuda,sass -k regex:<kernel> --launch-count 1 [--launch-skip N] > src.csv
    python scripts/ncu_hot_lines.py src.csv [top]

Per source line (file:line): share of the warp-stall samples, share of the executed warp instructions, the leading stall
reasons (of the not-issued samples), and the source text.
"""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    rows = list(csv.reader(open(path, newline="")))
    cur_file, hdr, fn = None, None, None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; hdr = None
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":          # a CUDA source line (SASS rows carry an address)
            d = dict(zip(hdr, r))
            lines.append((cur_file, int(r[0]), r[1], d))
    stall_cols = [c for c in hdr if c.startswith("stall_") and c.endswith("(Not Issued)")]
    tot_s = sum(float(d["# Samples"] or 0) for _, _, _, d in lines)
    tot_i = sum(float(d["Instructions Executed"] or 0) for _, _, _, d in lines)
    print("kernel: %s" % fn)
    print("samples %d, warp instructions %.4g, source lines with samples %d" % (tot_s, tot_i, sum(1 for l in lines if float(l[3]["# Samples"] or 0) > 0)))
    tot_st = defaultdict(float)
    for _, _, _, d in lines:
        for c in stall_cols:
            tot_st[c] += float(d[c] or 0)
    ns = sum(tot_st.values())
    print("not-issued samples by reason: " + ", ".join("%s %.0f%%" % (c[6:-13], 100 * v / ns) for c, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:7]))
    print()
    print("%-22s %7s %7s  %-34s %s" % ("file:line", "samp%", "inst%", "leading stalls", "source"))
    cum = 0.0
    for f, ln, src, d in sorted(lines, key=lambda l: -float(l[3]["# Samples"] or 0))[:top]:
        s = float(d["# Samples"] or 0)
        st = sorted(((float(d[c] or 0), c[6:-13]) for c in stall_cols), reverse=True)[:2]
        nst = sum(float(d[c] or 0) for c in stall_cols) or 1.0
        cum += s
        print("%-22s %6.1f%% %6.1f%%  %-34s %s" % ("%s:%d" % (f, ln), 100 * s / tot_s, 100 * float(d["Instructions Executed"] or 0) / tot_i,
                                                   ", ".join("%s %.0f%%" % (n, 100 * v / nst) for v, n in st if v > 0), " ".join(src.split())[:110]))
    print("\nthe %d lines above hold %.0f%% of the samples" % (top, 100 * cum / tot_s))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
