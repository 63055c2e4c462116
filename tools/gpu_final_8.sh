#!/bin/bash
# 8-GPU box, final code: multi-GPU tests, bench at N = 8 (default = coset partition, streaming all-gather) and the north-star's
# all-to-all partition for the record
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_gpu_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_8.log
tail -5 gpurun_out/pytest_gpu_8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/final_8gpu.json 2> gpurun_out/final_8gpu.err; echo "bench 8 rc=$?"
tail -c 300 gpurun_out/final_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --exchange alltoall --no-multi-in-process --large-commit off > gpurun_out/final_8gpu_alltoall.json 2> gpurun_out/final_8gpu_alltoall.err; echo "bench 8 alltoall rc=$?"
tail -c 600 gpurun_out/final_8gpu_alltoall.err
