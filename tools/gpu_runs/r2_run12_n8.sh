#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r2l_topo.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus 8 --steps 3 --warmup 2 --deadline 900 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err
tail -c 600 gpurun_out/r2l_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench_n8.json'))
print('N=8 headline', d.get('value'), d.get('e2e'), d.get('scaling'), d.get('config',{}).get('parallelism'), d.get('tp_check'))
print('roof', d['roofline']['whole_path'] if d.get('roofline') else None)
print({k:(v['launches_per_round'], round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()} if d.get('roofline') else None)
ex=d.get('extra',{})
print('replicas', ex.get('replicas'))
for k,v in (ex.get('tp') or {}).items():
    if isinstance(v, dict):
        print(k, {kk: vv for kk, vv in v.items() if kk != 'per_class'})
        print('   ', {kk:(vv['launches_per_round'], round(vv['ms_per_round'],3)) for kk,vv in v.get('per_class',{}).items()})
    else:
        print(k, v)
PY
