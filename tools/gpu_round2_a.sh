#!/bin/bash
# round 2, pass A (1 GPU): parity tests, bench (ours + reference arm), launch list
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench.err
echo skip ref
tail -c 600 gpurun_out/bench_ref.err
nproc; free -g | head -2
