#!/bin/bash
# round 2, batch v: GroupNorm apply unroll -- GN / stage / UniPose tests + gdino_stage line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gdino_model_gpu.py tests/test_unipose_gpu.py tests/test_gdino_heads_gpu.py tests/test_internimage_gpu.py -m gpu -q 2>&1 | tail -3
timeout 200 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_v.json 2>/dev/null
python - <<'P'
import json
d = json.load(open("gpurun_out/r2_bench_gdino_stage_v.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_breakdown"]["groupnorm"])
P
