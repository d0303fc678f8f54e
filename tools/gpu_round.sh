#!/bin/bash
# one full GPU visit: tests + bench + reference arm + ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_err.log | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 | tee gpurun_out/bench_ref.json
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fc_chain -s 3 -c 1 -f -o gpurun_out/prof_fc \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
fi
ls -la gpurun_out | head -30
echo "== batch 4M (ramp/tail amortised)"
timeout 300 python bench.py --batch 4194304 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_4m.json | cut -c1-260
./tools/gpu_models.sh
