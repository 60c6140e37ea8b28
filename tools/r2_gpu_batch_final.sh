#!/bin/bash
# round 2, final validation of HEAD: full GPU suite, smoke(), default bench line (driver settings), launch list + refreshed stage lines
mkdir -p gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default (driver settings)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_pair_forward_final.json 2>gpurun_out/final_def.err; tail -2 gpurun_out/final_def.err
echo "== bench gdino_stage / internimage_h"
timeout 300 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_final.json 2>gpurun_out/final_gd.err; tail -2 gpurun_out/final_gd.err
timeout 300 python bench.py --workload internimage_h --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_internimage_h_final.json 2>gpurun_out/final_ii.err; tail -2 gpurun_out/final_ii.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*_final.json')):
    try:
        d = json.load(open(f))
        print(f, d['value'], d['unit'], 'ms', round(d['ms_per_step'], 2), 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'], d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
P
echo "== ncu launch list gdino_stage"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_gdino_final.csv python bench.py --workload gdino_stage --steps 1 --warmup 3 --cuprof --no-cpu-baseline > gpurun_out/final_nl.log 2>&1; tail -1 gpurun_out/final_nl.log | cut -c1-160; wc -l gpurun_out/r2_launches_gdino_final.csv
