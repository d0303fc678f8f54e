#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gather" 2>&1 | tail -4 | tee gpurun_out/v7_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs \
    2> gpurun_out/v7_bench2_err.log > gpurun_out/v7_bench2.json
tail -5 gpurun_out/v7_bench2_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/v7_bench2.json'))
print(json.dumps(d.get('gather'), indent=1))
print('value', d['value'], 'e2e', d['e2e']['value'], d['config'].get('host_numa'))
PY
