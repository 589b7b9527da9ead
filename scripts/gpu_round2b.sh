#!/bin/bash
# Round-2 second evidence visit: launch list of the training step as it is now, ncu --set full of the kernels added or changed
# since the first visit (march, render backward, sample compaction, training forward with the 96-wide head input).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 700 --csv --log-file gpurun_out/r2b_launches_train.csv env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
F="--set full --clock-control none --import-source on -f"
ncu $F -k regex:traverse_kernel -s 4 -c 1 -o gpurun_out/r2b_prof_traverse env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:render_bwd_kernel -s 4 -c 1 -o gpurun_out/r2b_prof_renderbwd env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:compact_samples_kernel -s 4 -c 1 -o gpurun_out/r2b_prof_compact env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k "regex:field_fwd_kernel<false, true" -s 4 -c 1 -o gpurun_out/r2b_prof_fieldsave env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:grid_bwd_kernel -s 8 -c 1 -o gpurun_out/r2b_prof_k2 env AHEAD=0 PROFILE=0 python scripts/train_profile.py > /dev/null 2>&1
ls -la gpurun_out/r2b_*
