#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  echo "== llm_tp x8 $tag"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --workload llm_tp --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_llm_tp_8_$tag.json 2>gpurun_out/tp8.err
  python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_bench_llm_tp_8_$tag.json')); print('$tag', d['value'], d['ms_per_step'])
except Exception as e: print('ERR', e)
P
  tail -2 gpurun_out/tp8.err
}
run c120_s28 VLLM_TP_COMPUTE_SMS=120 VLLM_TP_SCATTER_SMS=28
run c132_s16 VLLM_TP_COMPUTE_SMS=132 VLLM_TP_SCATTER_SMS=16
run c108_s40 VLLM_TP_COMPUTE_SMS=108 VLLM_TP_SCATTER_SMS=40
run c148_s28 VLLM_TP_COMPUTE_SMS=147 VLLM_TP_SCATTER_SMS=28
