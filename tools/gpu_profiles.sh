#!/bin/bash
# round-2 evidence visit (1 GPU): sweep, ncu launch list of the bench, full captures of the two dominant kernels, microbenchmarks
mkdir -p gpurun_out
timeout 900 python tools/sweep.py 2>/dev/null | tee gpurun_out/sweep_1gpu.md | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --sustain-seconds 0.05 > gpurun_out/r2_ncu_launch_run.log 2>&1
tail -2 gpurun_out/r2_ncu_launch_run.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_chain_kernel -s 6 -c 1 -f -o gpurun_out/prof_fc \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-latency --sustain-seconds 0.05 > gpurun_out/r2_ncu_fc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cnn_frontend16_tc -s 2 -c 1 -f -o gpurun_out/prof_cnn_tc \
    python bench.py --model cnn --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-latency --sustain-seconds 0.05 > gpurun_out/r2_ncu_cnn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fc_chain_kernel -s 6 -c 1 -f -o gpurun_out/prof_binary160 \
    python bench.py --model binary160 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-latency --sustain-seconds 0.05 > gpurun_out/r2_ncu_binary160.log 2>&1
./tools/bin/cnn_tail_bench > gpurun_out/r2_micro_cnn_tail.txt 2>&1
./tools/bin/tmem_ld_bench > gpurun_out/r2_micro_tmem_ld.txt 2>&1
./tools/bin/alu_rates > gpurun_out/r2_micro_alu_rates.txt 2>&1
ls -la gpurun_out | tail -12
