#!/bin/bash
# round 2, batch t: live-tile attention for UniPose's keypoint decoder mask
mkdir -p gpurun_out
echo "== gpu tests (attention, unipose)"; timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_unipose_gpu.py tests/test_swin_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/t_tests.log
echo "== bench unipose_stage"; timeout 300 python bench.py --workload unipose_stage --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_unipose_stage_t.json 2>gpurun_out/t_up.err; tail -3 gpurun_out/t_up.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r2_bench_unipose_stage_t.json'))
    print(d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], d['clocks'])
    print({k: (round(v['ms'], 2), v['launches']) for k, v in d.get('kernel_breakdown', {}).items()})
except Exception as e:
    print('ERR', e)
P
echo "== torch profile unipose_stage"; timeout 300 python tools/torch_profile.py unipose_stage gpurun_out/r2_unipose_stage_torch_profile_t.json > gpurun_out/t_prof.log 2>&1; sed -n 3,24p gpurun_out/t_prof.log; grep -n "aten ops" -A 16 gpurun_out/t_prof.log
