#!/bin/bash
# second-tier GPU visit: new tests, MNIST-like distribution, encoding x batch sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for dist in uniform mnist; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --dist $dist 2>/dev/null | tee gpurun_out/bench_dist_$dist.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('$dist: value %.3f G img/s  ms/step %.4f  frac %.3f  sm_mhz %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))
"
done
timeout 1200 python tools/sweep.py 2>&1 | tee gpurun_out/sweep.md | tail -20
