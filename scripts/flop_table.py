#!/usr/bin/env python
"""The frozen operation-count table behind `roofline.achieved`: FP64 operations per primitive quartet of every integral class
as the engine counts them (flops_prim_quartet in csrc/vb_engine.cu: add / mul = 1, FMA = 2, Boys evaluation and the reciprocal
square root at their executed counts), general and asymptotic (T >= 40) regime, and the closed far-field form of vb_far.cuh.

    python scripts/flop_table.py > profiles/r2_flop_model_table.txt

The function body is taken from the engine source at run time (it is not exported), compiled on the host with vb_eri.cuh."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "valence_b200", "csrc")


def main():
    src = open(os.path.join(CSRC, "vb_engine.cu")).read()
    m = re.search(r"double flops_prim_quartet\(int tb, int tk, int far = 0\)\n\{.*?\n\}\n", src, re.S)
    assert m, "flops_prim_quartet not found"
    prog = '#include <cstdio>\n#include "vb_far.cuh"\nusing namespace vb;\n' + m.group(0) + r'''
int main()
{
    const char* nm[NPTYPE] = {"ss", "ps", "pp", "ds", "dp", "dd"};
    std::printf("class      general  asymptotic  far-field form\n");
    for (int a = 0; a < NPTYPE; ++a)
        for (int b = 0; b < NPTYPE; ++b) {
            std::printf("(%s|%s)  %9.0f", nm[a], nm[b], flops_prim_quartet(a, b));
            if (a < 3 && b < 3) std::printf("  %10.0f", flops_prim_quartet(a, b, 1)); else std::printf("  %10s", "-");
            if (a < 2 && b < 2) std::printf("  %14.0f", far_flops(a, b)); else std::printf("  %14s", "-");
            std::printf("\n");
        }
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.cpp")
        open(p, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", CSRC, "-o", exe, p])
        sys.stdout.write(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)


if __name__ == "__main__":
    main()
