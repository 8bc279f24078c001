#!/bin/bash
mkdir -p gpurun_out
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee gpurun_out/r2s_round.log
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -x --timeout 600 2>&1 | tail -2 | tee -a gpurun_out/r2s_round.log
