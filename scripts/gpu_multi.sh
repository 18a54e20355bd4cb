#!/bin/bash
# Multi-GPU checks (run under `gpurun --gpus N`): the torchrun-free C++ program and the torchrun parity script.
N=${1:-2}
mkdir -p gpurun_out
timeout 300 build/dist_test $N > gpurun_out/dist_test_$N.log 2>&1; echo "[dist_test world=$N] rc=$? $(tail -1 gpurun_out/dist_test_$N.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_gpu_check_$N.log 2>&1
echo "[multi_gpu_check world=$N] rc=$? $(grep MULTI_GPU_CHECK gpurun_out/multi_gpu_check_$N.log)"
