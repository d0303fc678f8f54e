#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8 | tee gpurun_out/v5_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs \
    2> gpurun_out/v5_bench2_err.log | tee gpurun_out/v5_bench2.json | cut -c1-300
tail -15 gpurun_out/v5_bench2_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/v5_bench2.json'))
print(json.dumps(d.get('gather'), indent=1))
print('value', d['value'], 'e2e', d['e2e'])
PY
