#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cnn_tc_debug.py 2 18 2>&1 | tail -2 | tee gpurun_out/v4_cnn_fe2.log
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/v4_pytest_gpu.log
