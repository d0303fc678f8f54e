#!/bin/bash
# round-2 GPU visit: CNN front-end check per kernel (own process, bounded), tests, bench, reference arm, latency knobs, ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -2 | tee gpurun_out/v_gpu.txt
for fe in 1 2; do timeout 300 python tools/cnn_tc_debug.py $fe 18 2>&1 | tail -8 | tee gpurun_out/v_cnn_fe$fe.log; done
if grep -q "logits False\|Error\|error" gpurun_out/v_cnn_fe2.log || ! grep -q "G images/s" gpurun_out/v_cnn_fe2.log; then
  echo "tensor-core CNN front-end NOT validated: tests run with the CUDA-core kernel as default" | tee -a gpurun_out/v_cnn_fe2.log
  export BNM_CNN_FRONTEND=1
fi
timeout 1700 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/v_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/v_bench_err.log | tee gpurun_out/v_bench.json | cut -c1-400
tail -12 gpurun_out/v_bench_err.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tee gpurun_out/v_bench_ref.json | cut -c1-300
for zc in 0 1 2; do
  BNM_SMALL_ZC=$zc timeout 120 python -c "import bench, json; print(json.dumps(bench.latency_batch1('fc', 20000)))" 2>&1 | tail -1 | tee gpurun_out/v_latency_zc$zc.json
done
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cnn_frontend16_tc -s 2 -c 1 -f -o gpurun_out/prof_cnn_tc \
      python tools/cnn_tc_debug.py 2 17 > gpurun_out/v_ncu_cnn.log 2>&1
  tail -3 gpurun_out/v_ncu_cnn.log
fi
ls -la gpurun_out | tail -15
