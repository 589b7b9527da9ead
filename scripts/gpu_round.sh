#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu --set full of the top kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc > gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-codec --train-steps 0 --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
ncu --set full --clock-control none --import-source on -k regex:field_fwd_kernel -s 3 -c 1 -f -o gpurun_out/prof_field \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-codec --train-steps 0 --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:context3d_kernel -s 19 -c 1 -f -o gpurun_out/prof_context \
    python scripts/codec_time.py > gpurun_out/ncu_context.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ac_decode_kernel -s 5 -c 1 -f -o gpurun_out/prof_decode \
    python scripts/codec_time.py > gpurun_out/ncu_decode.log 2>&1
