#!/bin/bash
# 2-GPU pass: real-NCCL tests, in-process multi-GPU tests, bench at N=2 (ours) 
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_2.log
tail -30 gpurun_out/pytest_gpu_2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/bench_2gpu.err
