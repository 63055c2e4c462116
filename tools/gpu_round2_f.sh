#!/bin/bash
# 8-GPU pass: real-NCCL tests at world 2/4/8, in-process multi-GPU tests, bench at N = 8 and 4
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nproc
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_gpu_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_8.log
tail -12 gpurun_out/pytest_gpu_8.log
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
tail -c 600 gpurun_out/bench_${N}gpu.err
done
