#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2w_sanity.log
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -x --timeout 600 2>&1 | tail -2 | tee -a gpurun_out/r2w_sanity.log
timeout 200 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -1 | tee -a gpurun_out/r2w_sanity.log
