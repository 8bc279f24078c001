#!/bin/bash
# compute-sanitizer over a tiny end-to-end run (both prefill paths, greedy + sampled rounds, AR).
mkdir -p gpurun_out
export LSK_SANITIZE=1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_probe.py > gpurun_out/r2p_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2p_memcheck.log
tail -4 gpurun_out/r2p_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_probe.py > gpurun_out/r2p_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2p_racecheck.log
tail -4 gpurun_out/r2p_racecheck.log
