#!/bin/bash
mkdir -p gpurun_out
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee gpurun_out/r2q_round.log
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x --timeout 600 2>&1 | tail -2 | tee -a gpurun_out/r2q_round.log
for sp in 8 2; do LSK_ATTN_SPLITS=$sp timeout 200 python tools/profile_round.py llama3-8b 20 400 2>&1 | tail -1 | tee -a gpurun_out/r2q_round.log; done
export LSK_PREFILL_TC=0
timeout 600 ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attn_cluster" -s 5132 -c 5 -o gpurun_out/r2_full_verify_b python tools/profile_round.py llama2-7b 1 400 > gpurun_out/r2q_ncu.log 2>&1
ncu -i gpurun_out/r2_full_verify_b.ncu-rep --page raw --csv > gpurun_out/r2_full_verify_b.raw.csv 2>/dev/null
rm -f gpurun_out/r2_full_verify_b.ncu-rep
