#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 4 --steps 2 --warmup 1 --deadline 500 > gpurun_out/r2u_bench_n4.json 2> gpurun_out/r2u_bench_n4.err
tail -c 300 gpurun_out/r2u_bench_n4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_n4.json'))
print('N=4 headline', d.get('value'), d.get('e2e',{}).get('value'), d.get('tp_check'))
print('whole', d['roofline']['whole_path'] if d.get('roofline') else None)
ex=d.get('extra',{})
print('replicas', ex.get('replicas'))
for k,v in (ex.get('tp') or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk != 'per_class'} if isinstance(v, dict) else v)
PY
