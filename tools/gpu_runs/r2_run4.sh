#!/bin/bash
mkdir -p gpurun_out
echo "== shapes, decode-kernel prefill" | tee gpurun_out/r2d_shapes.log
LSK_PREFILL_TC=0 timeout 1200 python -m pytest tests/test_gpu_shapes.py -q -s --timeout 600 2>&1 | grep -E "ctx=|worst|passed|failed|Error|assert" | tee -a gpurun_out/r2d_shapes.log
echo "== tcgen05 prefill unit" | tee -a gpurun_out/r2d_shapes.log
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x --timeout 300 -k "tcgen05_prefill or long or golden" 2>&1 | tail -15 | tee -a gpurun_out/r2d_shapes.log
echo "== shapes, tcgen05 prefill" | tee -a gpurun_out/r2d_shapes.log
timeout 1200 python -m pytest tests/test_gpu_shapes.py -q -s --timeout 600 2>&1 | grep -E "ctx=|worst|passed|failed|Error|assert" | tee -a gpurun_out/r2d_shapes.log
for sp in 8 4 2 1; do for st in 2 4; do
  LSK_ATTN_SPLITS=$sp LSK_ATTN_STAGES=$st timeout 200 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -1 | tee -a gpurun_out/r2d_attn_sweep.log
done; done
LSK_ATTN_SPLITS=2 LSK_ATTN_STAGES=4 timeout 200 python tools/profile_round.py llama3-8b 20 400 2>&1 | tail -1 | tee -a gpurun_out/r2d_attn_sweep.log
LSK_ATTN_SPLITS=8 LSK_ATTN_STAGES=2 timeout 200 python tools/profile_round.py llama3-8b 20 400 2>&1 | tail -1 | tee -a gpurun_out/r2d_attn_sweep.log
