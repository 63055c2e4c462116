#!/bin/bash
set -x
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "multi_commit or c_abi" > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$N.log
tail -5 gpurun_out/pytest_multi_$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --large-commit off > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_${N}gpu.err
