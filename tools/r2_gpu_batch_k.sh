#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_gemm_k256 python tools/ncu_k256.py 2>&1 | tail -5
