#!/bin/bash
# Everything that was written after the round-1 GPU budget ran out, in one gpurun call:
#   gpurun --gpus 2 --timeout 1800 -- 'bash tools/unrun_checks.sh'
# Each step runs under its own timeout so a hang cannot eat the box; logs land in gpurun_out/.
mkdir -p gpurun_out
set -x
# 1. checkpoint-fed engine == state-dict-fed engine (1 GPU)
timeout 300 python -m pytest tests/test_gpu_zz_checkpoint.py -x -q 2>&1 | tail -5 | tee gpurun_out/unrun_checkpoint.log
# 1b. push-merge attention kernel == pull-merge kernel, and what it buys on the 7B round
LSK_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_experimental.py -q -s 2>&1 | tail -5 | tee gpurun_out/unrun_attn_push.log
for v in 0 1; do LSK_ATTN_PUSH=$v timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee -a gpurun_out/unrun_attn_push.log; done
LSK_LMHEAD_TC=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee -a gpurun_out/unrun_attn_push.log
# 2. one-shot peer collectives == NCCL path, bit for bit at TP=2; sampling under TP (2 GPUs)
LSK_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_tp.py -x -q 2>&1 | tail -8 | tee gpurun_out/unrun_tp_oneshot.log
# 3. what it buys: 13B at TP=2, NCCL vs one-shot (short runs)
for mode in 0 1 2; do
  LSK_TP_ONESHOT=$mode timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --tp --arch llama2-13b \
    --steps 2 --warmup 3 --max-steps 128 --no-extra --no-cpu-baseline > gpurun_out/unrun_tp13b_oneshot$mode.json 2> gpurun_out/unrun_tp13b_oneshot$mode.err
  tail -c 600 gpurun_out/unrun_tp13b_oneshot$mode.json
done
