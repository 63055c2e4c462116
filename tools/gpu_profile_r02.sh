#!/bin/bash
# round 2: plain bench, then (same command, nothing changed) the ncu launch list and --set full captures of the leaf-hash
# kernel and of the two NTT passes.  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fri --no-from-values"
$CMD > gpurun_out/plain_prof.json 2> gpurun_out/plain_prof.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_prof2.json 2> gpurun_out/plain_prof2.err && \
ncu --set full --clock-control none --import-source on -k regex:k_hash_cols -s 3 -c 1 -o gpurun_out/r02_prof_leaf $CMD > gpurun_out/ncu_f.log 2>&1
echo "full leaf rc=$?"
$CMD > gpurun_out/plain_prof3.json 2> gpurun_out/plain_prof3.err && \
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 6 -c 2 -o gpurun_out/r02_prof_ntt $CMD > gpurun_out/ncu_n.log 2>&1
echo "full ntt rc=$?"
ls -la gpurun_out/*.ncu-rep
