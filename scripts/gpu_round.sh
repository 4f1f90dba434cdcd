#!/bin/bash
# Round-end style GPU pass: parity tests, bench, ncu launch list of the bench command, DRAM traffic of the hot kernel.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --cpu-baseline-seconds 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_ptile --launch-skip 1 --launch-count 1 \
    --csv --log-file gpurun_out/dram_k_ptile_H2O256.csv python scripts/ncu_tile.py 256 1 > gpurun_out/dram256.log 2>&1
VB_DEBUG_TIME=1 python scripts/ncu_tile.py 256 2 2>&1 | tail -6
