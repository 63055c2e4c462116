#!/bin/bash
# N-GPU box: the real-NCCL tests and the in-process multi-GPU tests (everything that needs more than one GPU)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -8 gpurun_out/pytest_gpu_multi.log
