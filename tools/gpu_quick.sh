#!/bin/bash
# quick GPU visit: parity tests + bench (no CPU baseline) [+ optional ncu full capture with NCU=1]
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} 2> gpurun_out/bench_err.log | tee gpurun_out/bench.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  frac %.3f  kernel_ms %.4f  e2e %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e'] and '%.1f M' % (d['e2e']['value']/1e6)))
"
tail -3 gpurun_out/bench_err.log
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_chain -s 3 -c 1 -f -o gpurun_out/prof_fc \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_full_run.log 2>&1
fi
