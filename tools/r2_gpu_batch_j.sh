#!/bin/bash
mkdir -p gpurun_out
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2_gputests_j.log | tail -30
