#!/bin/bash
mkdir -p gpurun_out
echo "== ncu launch list pair_forward"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_pair.csv python bench.py --steps 1 --warmup 3 --cuprof --no-cpu-baseline > gpurun_out/nl1.log 2>&1; tail -2 gpurun_out/nl1.log | cut -c1-200; wc -l gpurun_out/r2_launches_pair.csv
echo "== ncu launch list gdino_stage"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_gdino.csv python bench.py --workload gdino_stage --steps 1 --warmup 3 --cuprof --no-cpu-baseline > gpurun_out/nl2.log 2>&1; tail -2 gpurun_out/nl2.log | cut -c1-200; wc -l gpurun_out/r2_launches_gdino.csv
for wl in pair_forward_gdino cfg1_forward internimage_h; do
echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_${wl}_n.json 2>gpurun_out/$wl.err; python - <<P
import json
try:
    d=json.load(open('gpurun_out/r2_bench_${wl}_n.json')); print(d['value'], d['unit'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print('ERR', e)
P
tail -2 gpurun_out/$wl.err
done
