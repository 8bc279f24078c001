#!/bin/bash
# ncu --set full captures of the shipped decode kernels (7B shapes, steady state) and of the tcgen05
# prefill kernels.  The decode capture runs WITHOUT the second (prefill) weight copy so that ncu's
# per-kernel save / restore of device memory stays at round-1 size.
mkdir -p gpurun_out
export LSK_PREFILL_TC=0
# prefill (decode-kernel path, 400 ids = 25 chunks x 32 layers x 5 kernels = 4000 launches) + 2 warm rounds
# (2 x 407 matching launches): start inside round 3 -> draft step (M = 1) kernels
timeout 900 ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attn_cluster" -s 4830 -c 10 -o gpurun_out/r2_full_draft python tools/profile_round.py llama2-7b 1 400 > gpurun_out/r2m_ncu1.log 2>&1
tail -2 gpurun_out/r2m_ncu1.log
timeout 900 ncu --set full --clock-control none -k regex:"gemm_skinny_kernel|attn_cluster" -s 5132 -c 10 -o gpurun_out/r2_full_verify python tools/profile_round.py llama2-7b 1 400 > gpurun_out/r2m_ncu2.log 2>&1
tail -2 gpurun_out/r2m_ncu2.log
unset LSK_PREFILL_TC
timeout 600 ncu --set full --clock-control none -k regex:"prefill_gemm_tc|rms_canon" -s 7 -c 7 -o gpurun_out/r2_full_prefill python tools/prefill_probe.py llama2-7b-l2 128 > gpurun_out/r2m_ncu3.log 2>&1
tail -2 gpurun_out/r2m_ncu3.log
for f in r2_full_draft r2_full_verify r2_full_prefill; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ls -la gpurun_out/$f.ncu-rep gpurun_out/$f.raw.csv
done
find gpurun_out -name "*.ncu-rep" -size +18M -delete
