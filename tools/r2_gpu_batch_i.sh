#!/bin/bash
mkdir -p gpurun_out
echo "== train tests"; timeout 900 python -m pytest tests/test_train_gpu.py tests/test_swin_gpu.py tests/test_gdino_model_gpu.py -q 2>&1 | tail -25
echo "== bench llm_train"; timeout 600 python bench.py --workload llm_train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_llm_train_i.json 2>gpurun_out/train.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_llm_train_i.json')); print(d['value'], d['ms_per_step']); print(d['kernel_breakdown'])
P
tail -5 gpurun_out/train.err
echo "== torch profile llm_train"; timeout 400 python tools/torch_profile.py llm_train gpurun_out/r2_llm_train_torch_profile_i.json 2>&1 | head -22
