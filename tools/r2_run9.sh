#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prefill_probe.py llama2-7b 128 1024 2>&1 | tail -2 | tee gpurun_out/r2i_prefill.log
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_engine.py -q -x -s --timeout 600 2>&1 | grep -E "acceptance:|passed|failed|Error" | tee gpurun_out/r2i_tests.log
# ---- ncu --set full: decode kernels in steady state (draft M=1, then verify M=7), tcgen05 prefill kernels
timeout 900 ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attn_cluster" -s 960 -c 10 -o gpurun_out/r2_full_draft python tools/profile_round.py llama2-7b 1 400 > gpurun_out/r2i_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attn_cluster" -s 1262 -c 10 -o gpurun_out/r2_full_verify python tools/profile_round.py llama2-7b 1 400 > gpurun_out/r2i_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"prefill_gemm_tc|rms_canon" -s 14 -c 7 -o gpurun_out/r2_full_prefill python tools/prefill_probe.py llama2-7b 128 > gpurun_out/r2i_ncu3.log 2>&1
for f in r2_full_draft r2_full_verify r2_full_prefill; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ls -la gpurun_out/$f.ncu-rep
done
# keep the reports small enough to travel
find gpurun_out -name "*.ncu-rep" -size +20M -delete
# ---- compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_probe.py > gpurun_out/r2i_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2i_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_probe.py > gpurun_out/r2i_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2i_racecheck.log
tail -5 gpurun_out/r2i_memcheck.log gpurun_out/r2i_racecheck.log
