#!/bin/bash
mkdir -p gpurun_out
BNM_TRACE=gpurun_out/trace_full.txt timeout 120 python - <<'PY'
import sys; sys.path.insert(0, '.')
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
import torch
m = Model.load('tests/golden/models/fc.bnm'); e = Engine(m)
n = 1 << 20
x = torch.randint(-128, 128, (n, 256), dtype=torch.int8, device='cuda')
lo = torch.empty((n, 10), dtype=torch.int32, device='cuda'); la = torch.empty(n, dtype=torch.int32, device='cuda')
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); e.infer_device(x, lo, la); b.record(); torch.cuda.synchronize()
    print('launch ms (events, trace build incl. sync dump)', a.elapsed_time(b))
PY
python tools/trace_report.py gpurun_out/trace_full.txt 4 2
