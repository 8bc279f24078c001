#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prefill_probe.py llama2-7b 128 1024 2>&1 | tail -3 | tee gpurun_out/r2g_prefill.log
timeout 600 python -m pytest tests/test_gpu_engine.py -q -x --timeout 300 -k "tcgen05_prefill or long or logits_close" 2>&1 | tail -5 | tee -a gpurun_out/r2g_prefill.log
timeout 900 python -m pytest tests/test_gpu_shapes.py -q -x --timeout 600 -k "w7b or l32" 2>&1 | tail -5 | tee -a gpurun_out/r2g_prefill.log
timeout 300 python tools/sampling_probe.py 256 2>&1 | tail -1 | tee gpurun_out/r2g_sampling.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prefill_gemm|rms_canon|attn_cluster|embed" -c 260 --csv --log-file gpurun_out/r2g_prefill_launches.csv python tools/prefill_probe.py llama2-7b 128 > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/r2g_prefill_launches.csv') if not l.startswith('=='))]
agg=collections.defaultdict(list)
for r in rows:
    if r.get('Metric Name')=='gpu__time_duration.sum':
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
        v*= {'ns':1e-3,'us':1.0,'ms':1e3,'nsecond':1e-3,'usecond':1.0,'msecond':1e3}.get(u,1.0)
        agg[r['Kernel Name'][:60]+' grid='+r.get('Grid Size','')].append(v)
for k,v in sorted(agg.items()):
    print(f"{k:100s} n={len(v):4d} mean={sum(v)/len(v):9.2f} us")
PY
