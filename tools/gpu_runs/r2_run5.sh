#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 -s 2>&1 | grep -vE "^  (w|l)[0-9a-z_]+ ctx=" | tail -40 | tee gpurun_out/r2e_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench_7b.json 2> gpurun_out/r2e_bench_7b.err
tail -c 1500 gpurun_out/r2e_bench_7b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_7b.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'whole',d['roofline']['whole_path']['frac'],'dom',d['roofline']['frac'])
print({k:(v['launches_per_round'],round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()})
print(d['extra'].get('prefill_ms'), d['extra'].get('autoregressive_same_engine'))
print(d['cpu_baseline'])
PY
