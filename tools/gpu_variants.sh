#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./tools/bin/requant_bench | grep -E "correct"; BIGROWS=1 timeout 60 ./tools/bin/requant_bench | tee gpurun_out/requant_bench.log | grep -E "^[0-9]A|3 warps"
for lib in variants/*.so; do
  echo "== $lib"
  BNM_LIB_PATH=$PWD/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_and_oracle or edge_batch or full_size" 2>&1 | tail -1
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
echo "== default lib: CNN tests + bench"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "cnn or kat or edge" 2>&1 | tail -1
MODELS="cnn cnn_48" ./tools/gpu_models.sh
