#!/bin/bash
# visit 2: split kernel (wide models) A/B, full tests, ncu of a LARGE CNN front-end launch and of the split kernel
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/v2_pytest_gpu.log
for sp in 1 0; do
  BNM_SPLIT=$sp timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-latency --steps 10 2> gpurun_out/v2_bench_split${sp}_err.log > gpurun_out/v2_bench_split$sp.json
  grep "^config" gpurun_out/v2_bench_split${sp}_err.log | sed "s/^/split=$sp /"
done
for sl in 1 2; do
  BNM_SLOTS=$sl timeout 300 python bench.py --model binary160 --no-e2e --no-cpu-baseline --no-latency --no-configs --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('binary160 slots=$sl', d['value']/1e9, d['roofline']['frac'])"
done
for sl in 1 2 4; do
  BNM_SLOTS=$sl timeout 300 python bench.py --model 2bitsym96 --no-e2e --no-cpu-baseline --no-latency --no-configs --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2bitsym96 slots=$sl', d['value']/1e9, d['roofline']['frac'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cnn_frontend16_tc -s 6 -c 1 -f -o gpurun_out/prof_cnn_tc \
    python tools/cnn_tc_debug.py 2 18 > gpurun_out/v2_ncu_cnn.log 2>&1
tail -2 gpurun_out/v2_ncu_cnn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_chain_split -s 4 -c 1 -f -o gpurun_out/prof_split \
    python bench.py --model binary160 --no-e2e --no-cpu-baseline --no-latency --no-configs --steps 3 > gpurun_out/v2_ncu_split.log 2>&1
tail -2 gpurun_out/v2_ncu_split.log
ls -la gpurun_out | tail -8
