#!/bin/bash
# Round-end style GPU pass on one B200: bench line, ncu launch list of the bench command, DRAM traffic of the hot kernel,
# full ncu capture of the contraction-only kernel of first_order_opt.
set -x
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 900 gpurun_out/bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --cpu-baseline-seconds 0 --grad-waters 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_ptile --launch-skip 1 --launch-count 1 \
    --csv --log-file gpurun_out/dram_k_ptile_H2O256.csv python scripts/ncu_tile.py 256 1 > gpurun_out/dram256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_contract --launch-skip 2 --launch-count 1 -f -o gpurun_out/r1_k_contract_n64 \
    python scripts/fo_time.py 64 > gpurun_out/ncu_contract.log 2>&1
tail -2 gpurun_out/ncu_contract.log
