#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
for cfg in ${CFGS:-"3 2" "3 1" "2 2"}; do
set -- ${cfg/_/ }
BNM_WG=$1 BNM_SLOTS=$2 BNM_TRACE=gpurun_out/trace_wg$1_s$2.txt timeout 120 python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, '.')
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
import torch
m = Model.load('tests/golden/models/fc.bnm'); e = Engine(m)
n = 1 << 20
x = torch.randint(-128, 128, (n, 256), dtype=torch.int8, device='cuda')
lo = torch.empty((n, 10), dtype=torch.int32, device='cuda'); la = torch.empty(n, dtype=torch.int32, device='cuda')
for _ in range(3): e.infer_device(x, lo, la)
torch.cuda.synchronize()
PY
echo "== WG=$1 SLOTS=$2"; python tools/trace_report.py gpurun_out/trace_wg$1_s$2.txt 4 $2
BNM_WG=$1 BNM_SLOTS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print('value %.3f G img/s  ms/step %.4f  frac %.3f  kernel_ms %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
"
done
