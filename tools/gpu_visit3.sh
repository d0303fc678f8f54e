#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cnn_tc_debug.py 2 18 2>&1 | tail -8 | tee gpurun_out/v3_cnn_fe2.log
./tools/bin/tmem_ld_bench 2>&1 | tee gpurun_out/v3_tmem_ld.txt
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/v3_pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cnn_frontend16_tc -s 6 -c 1 -f -o gpurun_out/prof_cnn_tc \
    python tools/cnn_tc_debug.py 2 18 > gpurun_out/v3_ncu_cnn.log 2>&1
tail -2 gpurun_out/v3_ncu_cnn.log
