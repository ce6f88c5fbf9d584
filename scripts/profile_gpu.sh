#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu captures of the top kernels.  Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
MATH=${1:-bf16}
B="python bench.py --math $MATH --no-cpu --no-accuracy --no-extras"
# 1) every launch of ~2 steps with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 420 --csv --log-file gpurun_out/launches_${MATH}.csv \
    $B --steps 3 --warmup 3 > gpurun_out/ncu_bench_${MATH}.log 2>&1
# 2) full-set captures of the dominant kernels (a few launches each, after the warm-up)
ncu --set full --clock-control none --import-source on -k regex:"tc_gather_gemm" -s 60 -c 6 -f -o gpurun_out/prof_tc_gemm \
    $B --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 30 -c 3 -f -o gpurun_out/prof_tc_wgrad \
    $B --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"warp_loss|head_dgrad_elu|head_tapsum|head_dpr9|adam|preprocess" -s 40 -c 8 -f -o gpurun_out/prof_hbm \
    $B --steps 1 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_${MATH}.csv
# 3) DRAM traffic of every tcgen05 launch of one step (cheap metrics only) -> gpurun_out/tc_traffic_${MATH}.csv
#    (scripts/summarize_profiles.py turns it into profiles/rNN_tc_traffic_<math>.json, which bench.py reports as roofline.traffic)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"tc_" -s 60 -c 60 --csv \
    --log-file gpurun_out/tc_traffic_${MATH}.csv python scripts/prof_heads.py $MATH > /dev/null 2>&1
