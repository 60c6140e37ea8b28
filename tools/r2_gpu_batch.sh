#!/bin/bash
# One gpurun call worth of round-2 evidence (1 GPU).  Everything lands under gpurun_out/ (scratch); summaries are copied to
# profiles/ afterwards.   usage: bash tools/r2_gpu_batch.sh [tag]
tag=${1:-a}
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -80 | tee gpurun_out/r2_gputests_$tag.log | tail -60
echo "== msda window sweep"; timeout 300 python tools/msda_win_sweep.py --out gpurun_out/r2_msda_window_sweep.json 2>&1 | tail -20
echo "== msda vs reference kernel"; timeout 200 python tools/msda_ref_bench.py --out gpurun_out/r2_msda_vs_reference_kernel.json > /dev/null 2>gpurun_out/msda_ref.err; tail -3 gpurun_out/msda_ref.err
echo "== bench cfg1"; timeout 300 python bench.py --workload cfg1_forward --steps 20 --warmup 5 > gpurun_out/r2_bench_cfg1_$tag.json 2>gpurun_out/cfg1.err; tail -c 600 gpurun_out/r2_bench_cfg1_$tag.json; tail -3 gpurun_out/cfg1.err
echo "== bench default"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_pair_forward_$tag.json 2>gpurun_out/pair.err; tail -c 1500 gpurun_out/r2_bench_pair_forward_$tag.json; tail -3 gpurun_out/pair.err
echo "== experimental FHFMA variants"; VLLM_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_msda_gpu.py tests/test_internimage_gpu.py -q -k fhfma 2>&1 | tail -5 | tee gpurun_out/r2_fhfma_$tag.log
echo "== bench gdino_stage"; timeout 400 python bench.py --workload gdino_stage --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_gdino_stage_$tag.json 2>gpurun_out/gd.err; tail -c 1200 gpurun_out/r2_bench_gdino_stage_$tag.json; tail -3 gpurun_out/gd.err
echo "== bench llm_train"; timeout 600 python bench.py --workload llm_train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_llm_train_$tag.json 2>gpurun_out/train.err; tail -c 1200 gpurun_out/r2_bench_llm_train_$tag.json; tail -5 gpurun_out/train.err
echo "== ncu targets (set full)"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_targets python tools/ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; tail -3 gpurun_out/ncu_targets.log
