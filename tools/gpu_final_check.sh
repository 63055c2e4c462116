#!/bin/bash
# last call of round 2: the full parity suite on the final build, then the ncu launch list of the bench command
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fri --no-from-values"
$CMD > gpurun_out/plain_final.json 2> gpurun_out/plain_final.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
