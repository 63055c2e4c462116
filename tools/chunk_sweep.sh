#!/bin/bash
# bench.py at N GPUs for several --chunks values (streaming coefficient exchange): device-resident and end-to-end ms
N=${1:-2}; shift
for c in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560 + c)) \
    bench.py --gpus $N --steps 5 --warmup 3 --chunks $c 2>/dev/null > gpurun_out/sweep_${N}_${c}.json
  python - "$c" gpurun_out/sweep_${N}_${c}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).readline())
print("chunks", sys.argv[1], "value ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 2),
      "LDE phase", round(d["phase_ms"]["FFT + blinding"], 2))
PY
done
