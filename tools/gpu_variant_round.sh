#!/bin/bash
# One GPU call: A/B the reduction forms (build/variants/*.so, tools/build_variants.py) with the reduced bench line (which also
# checks each variant against the oracle at the sample size), install the winner as the in-tree libpcs.so, run the whole
# `-m gpu` parity suite on it, then the full bench line.
mkdir -p gpurun_out
rm -f gpurun_out/ab_*.json gpurun_out/ab_*.err
python -c "import torch, numpy; print(torch.__version__)"   # page the image in once
FLAGS="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-fri --no-from-values"
for t in base d_n0 b_n2 h_n2; do
  PCS_LIB=$PWD/build/variants/libpcs_$t.so timeout 40 python bench.py $FLAGS > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err
  echo "ab $t rc=$? $(date +%s)"
done
python tools/variant_pick.py > gpurun_out/variant_choice.json; cat gpurun_out/variant_choice.json
timeout 75 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r3.log
tail -4 gpurun_out/pytest_r3.log
timeout 45 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r3.json
# instruction count and pipe shares of the installed build's leaf-hash kernel (a handful of metrics, one launch)
timeout 30 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_hash_cols -s 3 -c 1 --csv --log-file gpurun_out/r03_leaf_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fri --no-from-values > gpurun_out/r03_leaf_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r03_leaf_ncu.csv | cut -c1-400
