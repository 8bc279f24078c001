#!/bin/bash
# 2-GPU checks of the tensor-parallel paths (one-shot peer collectives vs NCCL)
mkdir -p gpurun_out
LSK_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_tp.py -x -q --timeout 600 2>&1 | tail -15 | tee gpurun_out/r2b_tp_tests.log
for arch in llama2-13b llama2-7b; do
for mode in 0 1 2; do
  LSK_TP_ONESHOT=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --tp --arch $arch \
    --steps 2 --warmup 3 --max-steps 128 --no-extra --no-cpu-baseline > gpurun_out/r2b_tp2_${arch}_mode$mode.json 2> gpurun_out/r2b_tp2_${arch}_mode$mode.err
  echo "$arch mode $mode: $(python -c "import json,sys; d=json.load(open('gpurun_out/r2b_tp2_${arch}_mode$mode.json')); print(d['value'], d['roofline']['whole_path']['frac'], {k:(v['launches_per_round'], round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/r2b_tp_bench.log
done; done
