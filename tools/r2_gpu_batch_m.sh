#!/bin/bash
mkdir -p gpurun_out
echo "== default bench x2 (driver-style launch)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_default_2.json 2>gpurun_out/def2.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_default_2.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'])
tp=d.get('tp',{}); print('tp', {k:tp.get(k) for k in ('value','ms_per_step','plain_schedule_ms_per_step','parity','tflops_per_gpu')}); print('train', tp.get('train'))
print('msda', {k:(v.get('ms'),v.get('frac')) if isinstance(v,dict) and 'ms' in v else None for k,v in d['msda'].items()})
P
tail -4 gpurun_out/def2.err
echo "== reference arm x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>gpurun_out/ref2.err | tail -c 700; tail -2 gpurun_out/ref2.err
