#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_chain -s 3 -c 1 -f -o gpurun_out/prof_fc \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log | cut -c1-200
