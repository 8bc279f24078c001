#!/bin/bash
# BASELINE.json configs[0] and configs[2] through the reference-shaped command lines
mkdir -p gpurun_out/cli
# config 0: correctness.py — 4-layer / 256-dim Llama, E=2, D=4, greedy, speculative vs autoregressive (token-exact)
timeout 300 python tools/correctness.py --model synthetic:survey-tiny --dataset synthetic --num_samples 8 \
  --exit_layer 2 --num_speculations 4 --max_steps 64 --sample False --model_args "alpha=0.1,max_ctx=512" \
  --output_dir gpurun_out/cli 2>&1 | tail -3 | tee gpurun_out/r2v_correctness_survey_tiny.log
# config 2: sweep.py — Llama-3-8B arch, exit_layer {4,8,12,16} x num_speculations {2,4,6,8}, alpha 0.1
timeout 1200 python tools/sweep.py --model synthetic:llama3-8b --dataset synthetic --num_samples 2 --max_steps 256 \
  --sample False --exit_layer_first 4 --exit_layer_last 16 --exit_layer_step 4 \
  --num_speculations_first 2 --num_speculations_last 8 --num_speculations_step 2 \
  --model_args "alpha=0.1,max_ctx=768" --output_dir gpurun_out/cli 2>&1 | grep "exit_layer" | tee gpurun_out/r2v_sweep_8b.log
cp gpurun_out/cli/sweep_*.csv gpurun_out/r2v_sweep_llama3_8b.csv 2>/dev/null
