#!/bin/bash
mkdir -p gpurun_out
for n in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n \
  tools/tp_logits_probe.py llama2-7b-l2 llama2-7b 2>&1 | grep "tp=" | tee -a gpurun_out/r2k_tp_logits.log
done
