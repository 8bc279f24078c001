#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -s -x --timeout 120 2>&1 | tail -25 | tee gpurun_out/r2c_attention.log
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py tests/test_gpu_sampling.py tests/test_gpu_zz_checkpoint.py -q -x --timeout 300 2>&1 | tail -15 | tee gpurun_out/r2c_engine.log
timeout 1500 python -m pytest tests/test_gpu_shapes.py -q -s -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/r2c_shapes.log
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -3 | tee gpurun_out/r2c_round.log
