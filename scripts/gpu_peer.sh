#!/bin/bash
# peer-memory exchange: tests, then the training step under both exchanges (N = number of GPUs of the box)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/peer_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/peer_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/peer_tests.txt
tail -5 gpurun_out/peer_tests.txt
for ex in nccl peer; do
  CNC_EXCHANGE=$ex PROFILE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/train_profile.py > gpurun_out/peer_step_${ex}_n$N.txt 2>&1
  grep -h "wall\|exchange:" gpurun_out/peer_step_${ex}_n$N.txt
done
CNC_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/train_profile.py > gpurun_out/peer_profile_n$N.txt 2>&1
grep -h "wall" gpurun_out/peer_profile_n$N.txt
