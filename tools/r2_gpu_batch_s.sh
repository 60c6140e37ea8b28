#!/bin/bash
# round 2, batch s: copy-free mask FPN, fp32 mask logits, UniPose attend-mask cache -- tests + the two stage benches
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/s_tests.log
echo "== bench gdino_stage"; timeout 300 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_s.json 2>gpurun_out/s_gd.err; tail -2 gpurun_out/s_gd.err
echo "== bench unipose_stage"; timeout 300 python bench.py --workload unipose_stage --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_unipose_stage_s.json 2>gpurun_out/s_up.err; tail -2 gpurun_out/s_up.err
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_gdino_stage_s.json', 'gpurun_out/r2_bench_unipose_stage_s.json'):
    try:
        d = json.load(open(f))
        print(f, d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], d['clocks'])
        print('   ', {k: (round(v['ms'], 2), v['launches']) for k, v in d.get('kernel_breakdown', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
P
echo "== torch profile gdino_stage (graph)"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_s.json > gpurun_out/s_prof.log 2>&1; sed -n 1,26p gpurun_out/s_prof.log
