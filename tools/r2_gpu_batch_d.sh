#!/bin/bash
# MSDA window kernel v2: tests, fill / geometry sweep, gdino_stage torch profile + bench
mkdir -p gpurun_out
echo "== msda tests"; timeout 600 python -m pytest tests/test_msda_window_gpu.py tests/test_msda_gpu.py tests/test_gdino_model_gpu.py tests/test_modules_gpu.py -q 2>&1 | tail -15
echo "== sweep"; timeout 600 python tools/msda_win_sweep.py --out gpurun_out/r2_msda_window_sweep_d.json 2>&1 | grep -o "'case.*" | cut -c1-170
echo "== torch profile gdino_stage"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile.json 2>&1 | tail -45
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_d.json 2>gpurun_out/gd.err; tail -c 800 gpurun_out/r2_bench_gdino_stage_d.json; tail -3 gpurun_out/gd.err
