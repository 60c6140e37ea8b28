#!/bin/bash
# round 2, batch u: UniPose stage under CUDA-graph replay
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_unipose_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --workload unipose_stage --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_unipose_stage_u.json 2>gpurun_out/u_up.err; tail -3 gpurun_out/u_up.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r2_bench_unipose_stage_u.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["launch"], d["gpu_launches"])
P
