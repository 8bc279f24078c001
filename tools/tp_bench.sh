#!/bin/bash
# usage: tools/tp_bench.sh N [extra bench args]  — runs bench.py on N GPUs of this box via torchrun
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
  --master-port 29517 bench.py --gpus "$N" "$@"
