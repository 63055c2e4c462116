#!/bin/bash
# 8-GPU box: the scaling series N = 1, 2, 4, 8 on ONE box (ours), plus the reference arm under torchrun (all-core check)
set -x
mkdir -p gpurun_out
for N in 1 2 4 8; do
if [ $N -eq 1 ]; then
python bench.py --gpus 1 --steps 10 --warmup 3 --no-fri > gpurun_out/scale_1gpu.json 2> gpurun_out/scale_1gpu.err; echo "bench 1 rc=$?"
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}gpu.json 2> gpurun_out/scale_${N}gpu.err; echo "bench $N rc=$?"
fi
tail -c 300 gpurun_out/scale_${N}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 --reference-full-size off > gpurun_out/scale_ref8.json 2> gpurun_out/scale_ref8.err; echo "ref rc=$?"
cat gpurun_out/scale_ref8.json | cut -c1-400
