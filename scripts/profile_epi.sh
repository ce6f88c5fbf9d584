#!/bin/bash
# Source-level ncu capture of a few NAMED layers (NVTX ranges of the engine, DOFB_NVTX=1): stall reasons per SASS line.
set -u
mkdir -p gpurun_out
INC=""
for t in "$@"; do INC="$INC --nvtx-include $t/"; done
DOFB_NVTX=1 ncu --set full --clock-control none --import-source on --nvtx $INC -c 6 -f -o gpurun_out/prof_epi python scripts/prof_heads.py bf16 > gpurun_out/prof_epi.log 2>&1
ls -la gpurun_out/prof_epi.ncu-rep
