#!/bin/bash
# launch-overlap modes side by side (BNM_OPT_LAUNCH_OVERLAP), headline workload
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
for ov in 2 1 0; do
  echo "== overlap $ov"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --overlap $ov ${BENCH_ARGS} 2> gpurun_out/bench_err.log | tee gpurun_out/bench_overlap$ov.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  frac %.3f  kernel_ms %.4f isolated %.4f e2e %s parity %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_isolated_launch'], d['e2e'] and '%.1f M' % (d['e2e']['value']/1e6), d['parity_vs_oracle_sample']))
"
  tail -3 gpurun_out/bench_err.log
done
