#!/bin/bash
mkdir -p gpurun_out
echo "== torch profile llm_train"; timeout 400 python tools/torch_profile.py llm_train gpurun_out/r2_llm_train_torch_profile.json 2>&1 | head -45
