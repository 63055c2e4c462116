#!/bin/bash
# NOTE: compute-sanitizer is CLOSED on this GPU pool (every call answers rc 86, "runs under it have left GPUs needing a reset"), so
# this script could not be used in round 2; memory safety of the new kernels rests on the bit-exact parity tests over ragged / tiny /
# odd-count shapes (tests/test_gpu_multi.py) instead.
# compute-sanitizer (ONE tool per gpurun call: $1 = memcheck | racecheck) on the plain-C smoke test: small tree (latency kernels,
# pinned small path), 2^16 leaves (throughput kernels, chunked host pipeline with the streaming sponge), from_values, pcs_multi (1 device)
set -x
TOOL=${1:-memcheck}
mkdir -p gpurun_out
gcc -O2 -I include tests/c_abi_smoke.c -o /tmp/c_abi_smoke -L plonky2_demo_b200 -lpcs -L oracle -loracle -Wl,-rpath,$PWD/plonky2_demo_b200 -Wl,-rpath,$PWD/oracle || exit 1
for LG in 4 13; do
  /tmp/c_abi_smoke 1 $LG > gpurun_out/sanitize_plain_$LG.log 2>&1 || { echo "plain run failed"; cat gpurun_out/sanitize_plain_$LG.log; exit 1; }
done
for LG in 4 13; do
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 /tmp/c_abi_smoke 1 $LG > gpurun_out/sanitize_${TOOL}_$LG.log 2>&1; echo "$TOOL lg_d=$LG rc=$?"
  tail -4 gpurun_out/sanitize_${TOOL}_$LG.log
done
