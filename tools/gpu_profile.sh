#!/bin/bash
# plain bench, then (same command, nothing changed) the ncu launch list and one --set full capture of the leaf-hash kernel
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_prof.json 2> gpurun_out/plain_prof.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v5.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_prof2.json 2> gpurun_out/plain_prof2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_hash_cols -s 3 -c 1 -o gpurun_out/prof_leaf_v5 $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc=$?"
