#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prefill_probe.py llama2-7b 128 1024 2>&1 | tail -2 | tee gpurun_out/r2i_prefill.log
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_engine.py -q -x -s --timeout 600 2>&1 | grep -E "acceptance:|passed|failed|Error" | tee gpurun_out/r2i_tests.log
