#!/bin/bash
# round 2, batch q: GDINO-stage glue kernels (row gather, sine embeddings, fused residual) -- tests, bench, aten-level profile
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/q_tests.log
echo "== bench gdino_stage"; timeout 300 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_q.json 2>gpurun_out/q_gd.err; tail -2 gpurun_out/q_gd.err
echo "== bench pair_forward_gdino"; timeout 400 python bench.py --workload pair_forward_gdino --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pair_forward_gdino_q.json 2>gpurun_out/q_pg.err; tail -2 gpurun_out/q_pg.err
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_gdino_stage_q.json', 'gpurun_out/r2_bench_pair_forward_gdino_q.json'):
    try:
        d = json.load(open(f))
        print(f, d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], d['clocks'])
        print('   ', {k: (round(v['ms'], 2), v['launches']) for k, v in d.get('kernel_breakdown', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
P
echo "== torch profile gdino_stage"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_q.json > gpurun_out/q_prof.log 2>&1; sed -n 1,30p gpurun_out/q_prof.log; grep -n "aten ops" -A 36 gpurun_out/q_prof.log
