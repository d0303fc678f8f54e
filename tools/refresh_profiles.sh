#!/bin/bash
# copy the evidence of the last tools/gpu_round.sh visit from gpurun_out/ (scratch) into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
R=${ROUND:-r1}
python tools/ncu_summary.py gpurun_out/prof_fc.ncu-rep 0 25 > profiles/${R}_fc_chain_ncu_summary.txt 2>&1
cp gpurun_out/launches.csv profiles/${R}_launches_bench.csv
cp gpurun_out/bench.json profiles/${R}_bench_1gpu.json
cp gpurun_out/bench_ref.json profiles/${R}_bench_reference_arm.json
python - <<PY
import json,re
names="cnn cnn_48 binary160 ternary64 12k_FP130 rand_fp130_64 1k 8bit64".split()
lines=open('gpurun_out/models.log').read().strip().splitlines()[-8:]
with open('profiles/${R}_models_sweep.txt','w') as f:
    f.write("# python bench.py --model <m> --steps 10 --warmup 3 --no-cpu-baseline --no-e2e   (batch 2^20, device-timed, B200, default launch-overlap mode 2)\n")
    for n,l in zip(names,lines): f.write(f"{n:14s} {l}\n")
txt=open('profiles/${R}_fc_chain_ncu_summary.txt').read()
rd=float(re.search(r'dram__bytes_read.sum\s+([\d.]+) Mbyte',txt).group(1)); wr=float(re.search(r'dram__bytes_write.sum\s+([\d.]+) Mbyte',txt).group(1))
json.dump({"fc:1048576": int(round((rd+wr)*1e6)), "note": "dram__bytes_read.sum + dram__bytes_write.sum of one fc_chain_kernel launch, ncu --set full (profiles/${R}_fc_chain_ncu_summary.txt)"}, open('profiles/traffic.json','w'))
print(open('profiles/traffic.json').read())
PY
head -16 profiles/${R}_fc_chain_ncu_summary.txt
