#!/bin/bash
# Round-2 evidence visit: ncu launch lists (forward bench step, training step, rate term, device renderer) and ncu --set full of
# the kernels that changed or were never captured.  Outputs under gpurun_out/; summaries are made here with scripts/ncu_summary.py.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-codec"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd.csv $B --train-steps 0 --no-e2e > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 900 --csv --log-file gpurun_out/r2_launches_train.csv python scripts/train_profile.py > /dev/null 2>&1
F="--set full --clock-control none --import-source on -f"
ncu $F -k regex:field_fwd_kernel -s 3 -c 1 -o gpurun_out/r2_prof_field $B --train-steps 0 --no-e2e > /dev/null 2>&1
ncu $F -k regex:traverse_kernel -s 4 -c 1 -o gpurun_out/r2_prof_traverse python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:adam_planes_kernel -s 4 -c 1 -o gpurun_out/r2_prof_adam python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:render_density_kernel -s 4 -c 1 -o gpurun_out/r2_prof_render python scripts/train_profile.py > /dev/null 2>&1
ncu $F -k regex:grid_fwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k1 python scripts/k1_time.py > /dev/null 2>&1
SAMPLE_NUM=150000 ncu $F -k regex:ctx3d_gather_fwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_ctxgather python scripts/rate_profile.py > /dev/null 2>&1
SAMPLE_NUM=150000 ncu $F -k regex:vote3_bwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_vote3bwd python scripts/rate_profile.py > /dev/null 2>&1
SAMPLE_NUM=150000 ncu $F -k regex:ctx_mlp_bwd_kernel -s 2 -c 1 -o gpurun_out/r2_prof_ctxmlpbwd python scripts/rate_profile.py > /dev/null 2>&1
ncu $F -k regex:query_mask_kernel -s 1 -c 1 -o gpurun_out/r2_prof_qmask python scripts/codec_time.py > /dev/null 2>&1
ncu $F -k regex:level_pruned_keys_kernel -s 17 -c 1 -o gpurun_out/r2_prof_prunedkeys python scripts/pruned_time.py > /dev/null 2>&1
ncu $F -k regex:wf_march_kernel -s 200 -c 1 -o gpurun_out/r2_prof_wfmarch python scripts/render_time.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -20
