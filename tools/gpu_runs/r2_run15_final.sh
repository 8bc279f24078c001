#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2o_smoke.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 | tee gpurun_out/r2o_pytest_gpu.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2o_bench_7b.json 2> gpurun_out/r2o_bench_7b.err
tail -c 500 gpurun_out/r2o_bench_7b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2o_bench_7b.json'))
print('value',d['value'],'e2e',d['e2e'],'whole',d['roofline']['whole_path']['frac'],'dom',d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
print({k:(v['launches_per_round'],round(v['ms_per_round'],3)) for k,v in d['roofline']['per_class'].items()})
print(json.dumps(d['extra'])[:1200])
print(d['cpu_baseline'])
PY
