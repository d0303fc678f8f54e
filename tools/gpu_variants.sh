#!/bin/bash
mkdir -p gpurun_out
for lib in variants/*.so; do
  echo "== $lib"
  for rep in 1 2; do
  BNM_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  frac %.3f  kernel_ms %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
" | tee -a gpurun_out/variants.log
  done
done
