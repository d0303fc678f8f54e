#!/bin/bash
# A/B of libraries (in-tree default + variants/*.so, + the default under BNM_STAGGER_NS values in STAGGERS) x launch-overlap
# modes OVS, interleaved
mkdir -p gpurun_out; : > gpurun_out/ab2.log
run() {  # label overlap
  timeout 300 python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --no-e2e --overlap $2 2>/dev/null | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('%-26s overlap %s  value %.3f G img/s  ms/step %.4f  frac %.3f  isolated %.4f  sm_mhz %s' % ('$1', '$2', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_isolated_launch'], d['clocks']['sm_mhz']))
" | tee -a gpurun_out/ab2.log
}
for rep in 1 2; do
  for ov in ${OVS:-0 2}; do
    unset BNM_LIB_PATH BNM_STAGGER_NS
    run default $ov
    for st in $STAGGERS; do export BNM_STAGGER_NS=$st; run "default stagger=$st" $ov; done
    unset BNM_STAGGER_NS
    for lib in $(ls variants/*.so 2>/dev/null); do export BNM_LIB_PATH=$PWD/$lib; run $lib $ov; done
  done
done
