#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_sharded.py -m gpu -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -25 gpurun_out/pytest_gpu_multi.log
