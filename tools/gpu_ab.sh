#!/bin/bash
# interleaved A/B of the in-tree library (optionally under several BNM_WG_BNM_SLOTS configs: CFGS="3_2 2_3") against
# variants/*.so -- power-capped boxes drift, never compare across visits
mkdir -p gpurun_out; : > gpurun_out/ab.log
run() {  # label
  timeout 300 python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('%-28s value %.3f G img/s  ms/step %.4f  frac %.3f  isolated %.4f  sm_mhz %s parity %s' % ('$1', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_isolated_launch'], d['clocks']['sm_mhz'], d['parity_vs_oracle_sample']))
" | tee -a gpurun_out/ab.log
}
for rep in 1 2 3; do
  for cfg in ${CFGS:-3_2}; do
    unset BNM_LIB_PATH; export BNM_WG=${cfg%_*} BNM_SLOTS=${cfg#*_}
    run "default wg${BNM_WG} slots${BNM_SLOTS}"
  done
  unset BNM_WG BNM_SLOTS
  for lib in $(ls variants/*.so 2>/dev/null); do
    export BNM_LIB_PATH=$PWD/$lib
    run "$lib"
  done
done
