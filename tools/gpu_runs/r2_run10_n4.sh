#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 4 --steps 3 --warmup 2 --deadline 700 > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err
tail -c 400 gpurun_out/r2j_bench_n4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_n4.json'))
print('N=4 headline', d.get('value'), d.get('scaling'), d.get('config',{}).get('parallelism'), d.get('tp_check'))
print('roof', d['roofline']['whole_path'] if d.get('roofline') else None)
print('extra', json.dumps(d.get('extra'))[:2500])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 4 --arch mini70b-tp4 --exit-layer 2 --steps 2 --warmup 1 --max-steps 256 --no-extra --deadline 500 > gpurun_out/r2j_mini70b_tp4.json 2> gpurun_out/r2j_mini70b_tp4.err
tail -c 400 gpurun_out/r2j_mini70b_tp4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_mini70b_tp4.json'))
print('mini70b tp4', d.get('value'), d.get('tp_check'), d['roofline']['whole_path'] if d.get('roofline') else None)
print({k:(v['launches_per_round'], round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()})
PY
