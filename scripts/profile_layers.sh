#!/bin/bash
# Full ncu captures of NAMED layers of one bf16 training step (B=32, 384x512): the engine wraps every launch in an NVTX range (DOFB_NVTX=1).
set -u
mkdir -p gpurun_out
INC=""
for t in "deconv_fwd:upconv1" "conv_dgrad:conv2" "conv_fwd:conv2" "conv_fwd:conv1" "conv_fwd:conv3_2" "conv_dgrad:conv4_2" "conv_wgrad:conv2" "conv_wgrad:conv1" \
         "conv_wgrad:conv4_2" "deconv_wgrad:upconv1" "elu_bwd:conv1" "elu_bwd:upconv1" "head_fwd:pr1" "head_wgrad:pr1" "head_dpr9:pr1" "warp_loss" "preprocess" "pack_weights" "adam"; do
  INC="$INC --nvtx-include $t/"
done
DOFB_NVTX=1 ncu --set full --clock-control none --nvtx $INC -f -o gpurun_out/prof_layers python scripts/prof_heads.py bf16 > gpurun_out/prof_layers.log 2>&1
# the report can exceed what travels back: export the raw page here, keep the report only when it is small
ncu -i gpurun_out/prof_layers.ncu-rep --page raw --csv --print-nvtx-rename kernel > gpurun_out/prof_layers_raw.csv 2>> gpurun_out/prof_layers.log
ls -la gpurun_out/prof_layers.ncu-rep gpurun_out/prof_layers_raw.csv
if [ $(stat -c %s gpurun_out/prof_layers.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/prof_layers.ncu-rep; fi
