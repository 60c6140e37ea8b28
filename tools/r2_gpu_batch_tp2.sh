#!/bin/bash
mkdir -p gpurun_out
echo "== tp tests (2 GPUs)"; timeout 600 python -m pytest tests/test_tp_gpu.py -q 2>&1 | tail -12
for wl in llm_tp llm_tp_plain; do
echo "== bench $wl x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_2.json 2>gpurun_out/tp2.err; python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_bench_${wl}_2.json')); print(d['value'], d['ms_per_step'], d['kernel_breakdown'])
except Exception as e: print('ERR', e)
P
tail -5 gpurun_out/tp2.err
done
