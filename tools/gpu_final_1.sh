#!/bin/bash
# 1 GPU, final code: what the driver runs at round end (parity suite, smoke, bench) + the reference arm
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_1gpu.json 2> gpurun_out/final_1gpu.err; echo "bench rc=$?"
tail -c 400 gpurun_out/final_1gpu.err
