#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== k256 gemm"; timeout 120 python tools/ncu_k256.py 2>&1 | grep "ms per launch"
echo "== gemm bench"; timeout 300 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench.json 2>gpurun_out/gb.err; tail -c 1500 gpurun_out/r2_gemm_bench.json; tail -2 gpurun_out/gb.err
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_l.json 2>gpurun_out/gd.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_gdino_stage_l.json')); print(d['value'], d['ms_per_step']); print(d['kernel_breakdown']['gemm'])
P
tail -3 gpurun_out/gd.err
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_pair_forward_l.json 2>gpurun_out/pair.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_pair_forward_l.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['frac'], d['roofline'].get('all_gemm_launches')); print(d['kernel_breakdown']); print(json.dumps(d['msda'])[:1200]); print(d['clocks'])
P
tail -3 gpurun_out/pair.err
