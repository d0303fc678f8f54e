#!/bin/bash
# multi-GPU evidence: bench line (with the fused gather figures) + quick encoding x batch sweep at N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-configs \
    2> gpurun_out/r2_bench_${N}gpu_err.log > gpurun_out/r2_bench_${N}gpu.json
tail -3 gpurun_out/r2_bench_${N}gpu_err.log
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
g=d.get('gather') or {}
print('N=$N value', d['value']/1e9, 'e2e', d['e2e']['value']/1e6, {k:(round(v/1e9,2) if isinstance(v,float) and v>1e6 else v) for k,v in g.items() if k.startswith('value') or k.endswith('correct') or 'error' in k})
PY
SWEEP_QUICK=${SWEEP_QUICK:-1} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/sweep.py 2>/dev/null | tee gpurun_out/sweep_${N}gpu.md | tail -4
