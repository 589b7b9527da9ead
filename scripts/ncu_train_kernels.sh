#!/bin/bash
# ncu --set full of the training-path kernels (one launch each), reports under gpurun_out/
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-codec --train-steps 1"
ncu --set full --clock-control none -k regex:wgrad_kernel -s 4 -c 1 -f -o gpurun_out/prof_wgrad $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:^dgrad_kernel -s 4 -c 1 -f -o gpurun_out/prof_dgrad $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:grid_bwd_kernel -s 4 -c 1 -f -o gpurun_out/prof_k2 $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:ac_encode_kernel -s 2 -c 1 -f -o gpurun_out/prof_encode python scripts/codec_time.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
