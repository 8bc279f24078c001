#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tp.py -x -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/r2t_tp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
  bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
tail -c 300 gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t_bench_n2.json'))
print('N=2 headline', d['value'], d['e2e']['value'], d['scaling'], d['config']['parallelism'], d.get('tp_check'))
print('whole', d['roofline']['whole_path'])
ex=d['extra']
print('replicas', ex.get('replicas'))
for k,v in (ex.get('tp') or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk != 'per_class'} if isinstance(v, dict) else v)
PY
