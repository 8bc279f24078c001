#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shapes.py -q -x -s --timeout 600 -k "w13b or w70b" 2>&1 | grep -E "m=9|m=8|worst|passed|failed|Error|assert" | tee gpurun_out/r2x_chunked_rms.log
timeout 300 python -m pytest tests/test_gpu_engine.py -q -x --timeout 300 -k "golden or equals" 2>&1 | tail -2 | tee -a gpurun_out/r2x_chunked_rms.log
