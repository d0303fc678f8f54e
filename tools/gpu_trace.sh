#!/bin/bash
mkdir -p gpurun_out
for wg in ${WGS:-4 5 6}; do
BNM_WG=$wg BNM_TRACE=gpurun_out/trace_wg$wg.txt timeout 120 python - <<'PY'
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
echo "== WG=$wg"; python tools/trace_report.py gpurun_out/trace_wg$wg.txt 4
done
true
