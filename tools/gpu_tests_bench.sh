#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/v6_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/v6_bench_err.log > gpurun_out/v6_bench.json
cat gpurun_out/v6_bench_err.log | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/v6_bench.json'))
print('value',d['value']/1e9,'ovl',d['value_overlapped_launches']/1e9,'sus',d['value_sustained']['value']/1e9,'frac',d['roofline']['frac'],'e2e',d['e2e']['value']/1e6,'lat',d['latency_us_batch1'], d['config'].get('host_numa'))
for c in d['configs']: print(c['name'], c.get('value',0)/1e9, c['roofline']['frac'] if 'roofline' in c else c, c.get('parity'), c.get('gpu_launches_per_step'), (c.get('e2e') or {}).get('value'))
PY
