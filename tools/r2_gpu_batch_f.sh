#!/bin/bash
mkdir -p gpurun_out
echo "== attention + unipose + gdino tests"; timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_unipose_gpu.py tests/test_gdino_model_gpu.py tests/test_modules_gpu.py tests/test_gemm_gpu.py tests/test_cfg1_e2e_gpu.py -q 2>&1 | tail -40
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_f.json 2>gpurun_out/gd.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_gdino_stage_f.json')); print(d['value'], d['ms_per_step']); print(d['kernel_breakdown'])
P
tail -3 gpurun_out/gd.err
echo "== torch profile gdino_stage"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_f.json 2>&1 | tail -32
