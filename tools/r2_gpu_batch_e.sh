#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_e.json 2>gpurun_out/gd.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_gdino_stage_e.json')); print(d['value'], d['ms_per_step'], d['kernel_breakdown'].get('msda_encoder'))
P
tail -3 gpurun_out/gd.err
echo "== ncu QP kernel in gdino_stage"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_win_kernel -c 1 -f -o gpurun_out/r2_msda_qp python bench.py --workload gdino_stage --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_qp.log 2>&1; tail -2 gpurun_out/ncu_qp.log
echo "== ncu targets"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_targets python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; tail -2 gpurun_out/ncu_targets.log
