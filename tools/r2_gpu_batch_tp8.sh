#!/bin/bash
mkdir -p gpurun_out
for wl in llm_tp llm_tp_plain; do
echo "== bench $wl x8"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_8.json 2>gpurun_out/tp8.err; python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_bench_${wl}_8.json')); print(d['value'], d['ms_per_step'], d['kernel_breakdown'])
except Exception as e: print('ERR', e)
P
tail -3 gpurun_out/tp8.err
done
