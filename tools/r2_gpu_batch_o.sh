#!/bin/bash
mkdir -p gpurun_out
echo "== swin / attention / gdino tests"; timeout 900 python -m pytest tests/test_swin_gpu.py tests/test_attention_gpu.py tests/test_gdino_model_gpu.py tests/test_cfg1_e2e_gpu.py -q 2>&1 | tail -12
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_o.json 2>gpurun_out/gd.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_gdino_stage_o.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_breakdown']['attention'])
P
tail -3 gpurun_out/gd.err
echo "== torch profile"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_o.json 2>&1 | head -14
