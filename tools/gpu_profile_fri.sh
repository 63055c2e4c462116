#!/bin/bash
# opening-proof kernels: plain run, then the ncu launch list and one --set full capture each of the two HBM-bound kernels
set -x
mkdir -p gpurun_out
CMD="python tools/fri_bench.py --steps 3"
$CMD > gpurun_out/fri_plain.json 2> gpurun_out/fri_plain.err || { tail -20 gpurun_out/fri_plain.err; exit 1; }
cat gpurun_out/fri_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fri_launches.csv $CMD > gpurun_out/fri_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_reduce_polys_base_tma -s 2 -c 1 -o gpurun_out/prof_fri_reduce $CMD > gpurun_out/fri_ncu_f1.log 2>&1
echo "full reduce rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_eval_ext_tma -s 2 -c 1 -o gpurun_out/prof_fri_eval $CMD > gpurun_out/fri_ncu_f2.log 2>&1
echo "full eval rc=$?"
