#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tp.py -x -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/r2h_tp_tests.log
for mode in 1 2 3 0; do
  LSK_TP_ONESHOT=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 --max-steps 256 --no-extra \
    > gpurun_out/r2h_tp2_7b_mode$mode.json 2> gpurun_out/r2h_tp2_7b_mode$mode.err
  echo "7b mode $mode: $(python -c "import json,sys; d=json.load(open('gpurun_out/r2h_tp2_7b_mode$mode.json')); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['whole_path']['frac'],3), d.get('tp_check'), {k:(v['launches_per_round'], round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()})" 2>&1 | tail -1)" | tee -a gpurun_out/r2h_tp_bench.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
tail -c 600 gpurun_out/r2h_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_n2.json'))
print('N=2 headline', d['value'], d['scaling'], d['config']['parallelism'], d.get('tp_check'))
print('extra', json.dumps(d['extra'])[:1500])
PY
