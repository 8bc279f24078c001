#!/bin/bash
mkdir -p gpurun_out
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee gpurun_out/r2n_round.log
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -x --timeout 600 2>&1 | tail -3 | tee -a gpurun_out/r2n_round.log
timeout 900 python -m pytest tests/test_gpu_shapes.py -q -x --timeout 600 -k "w13b or w70b or w7b" 2>&1 | tail -3 | tee -a gpurun_out/r2n_round.log
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama3-8b 20 400 2>&1 | tail -2 | tee -a gpurun_out/r2n_round.log
