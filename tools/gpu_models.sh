#!/bin/bash
mkdir -p gpurun_out
for m in ${MODELS:-cnn cnn_48 binary160 ternary64 12k_FP130 rand_fp130_64 1k 8bit64}; do
  echo "== $m"
  timeout 300 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  hbm-frac %.3f  kernel_ms %.4f launches %d path %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['config']['path']))
" | tee -a gpurun_out/models.log
done
