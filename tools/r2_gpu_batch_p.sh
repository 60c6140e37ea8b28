#!/bin/bash
# round 2, batch p: full GPU suite after the UniPose backbone / composite loss+pose / preset workloads; bench lines
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/p_tests.log
echo "== bench default"; timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_pair_forward_p.json 2>gpurun_out/p_def.err; tail -2 gpurun_out/p_def.err
echo "== bench gdino_stage"; timeout 300 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_p.json 2>gpurun_out/p_gd.err; tail -2 gpurun_out/p_gd.err
for w in pair_forward_clip7b pair_forward_clip7b_1tile pair_forward_1tile; do
  echo "== bench $w"; timeout 400 python bench.py --workload $w --steps 8 --warmup 3 > gpurun_out/r2_bench_${w}.json 2>gpurun_out/p_$w.err; tail -2 gpurun_out/p_$w.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*_p.json') + glob.glob('gpurun_out/r2_bench_pair_forward_*.json')):
    try:
        d = json.load(open(f))
        print(f, d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
        print('   ', {k: (round(v['ms'], 2), v['launches']) for k, v in d.get('kernel_breakdown', {}).items()})
    except Exception as e:
        print(f, 'ERR', e)
P
echo "== torch profile gdino_stage"; timeout 300 python tools/torch_profile.py gdino_stage gpurun_out/r2_gdino_stage_torch_profile_p.json 2>&1 | head -34
