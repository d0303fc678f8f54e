#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for wg in 6 5 4 3; do
  echo "== BNM_WG=$wg"
  BNM_WG=$wg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  frac %.3f  kernel_ms %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
" | tee -a gpurun_out/sweep.log
done
