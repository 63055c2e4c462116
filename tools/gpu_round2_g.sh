#!/bin/bash
# 1 GPU: full parity suite, bench, demo shapes
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench.err
python tests/harness/demo_shapes.py > gpurun_out/demo_shapes.json 2> gpurun_out/demo_shapes.err; echo "demo rc=$?"
tail -c 800 gpurun_out/demo_shapes.err
