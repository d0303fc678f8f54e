#!/bin/bash
# interleaved sweep of one environment knob:  KNOB=BNM_STAGGER_NS VALS="0 1000 2000" tools/gpu_knob.sh
mkdir -p gpurun_out; : > gpurun_out/knob.log
for rep in 1 2 3; do
  for v in $VALS; do
    export $KNOB=$v
    timeout 300 python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('%-24s value %.3f G img/s  ms/step %.4f  frac %.3f  isolated %.4f  sm_mhz %s' % ('$KNOB=$v', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_isolated_launch'], d['clocks']['sm_mhz']))
" | tee -a gpurun_out/knob.log
  done
done
