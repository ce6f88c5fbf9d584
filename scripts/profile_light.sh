#!/bin/bash
# Light refresh of the ncu evidence: launch list of ~2 steps + DRAM traffic of every tc_* launch of one step (no --set full captures).
set -u
mkdir -p gpurun_out
MATH=${1:-bf16}
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/launches_${MATH}.csv \
    python bench.py --math $MATH --no-cpu --no-accuracy --no-extras --steps 3 --warmup 3 > gpurun_out/ncu_bench_${MATH}.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"tc_" -s 60 -c 60 --csv \
    --log-file gpurun_out/tc_traffic_${MATH}.csv python scripts/prof_heads.py $MATH > /dev/null 2>&1
ls -la gpurun_out/launches_${MATH}.csv gpurun_out/tc_traffic_${MATH}.csv
