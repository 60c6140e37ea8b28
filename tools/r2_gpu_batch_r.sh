#!/bin/bash
# round 2, batch r: eager aten-level attribution of the GDINO stage's remaining torch glue; UniPose stage at real size
mkdir -p gpurun_out
echo "== eager torch profile gdino_stage"; VLLM_BENCH_GRAPH=0 timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_eager_r.json > gpurun_out/r_prof.log 2>&1; grep -n "aten ops" -A 42 gpurun_out/r_prof.log
echo "== bench unipose_stage"; timeout 400 python bench.py --workload unipose_stage --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_unipose_stage.json 2>gpurun_out/r_up.err; tail -12 gpurun_out/r_up.err
python - <<'P'
import json
try:
    d = json.load(open('gpurun_out/r2_bench_unipose_stage.json'))
    print(d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], d['clocks'], d['gpu_launches'])
    print({k: (round(v['ms'], 2), v['launches']) for k, v in d.get('kernel_breakdown', {}).items()})
except Exception as e:
    print('ERR', e)
P
echo "== torch profile unipose_stage"; timeout 300 python tools/torch_profile.py unipose_stage gpurun_out/r2_unipose_stage_torch_profile.json > gpurun_out/r_prof_up.log 2>&1; sed -n 1,28p gpurun_out/r_prof_up.log; grep -n "aten ops" -A 25 gpurun_out/r_prof_up.log
