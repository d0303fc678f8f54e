"""Launch-overlap modes of the fused FC kernel, interleaved in one process: us per step at batch 2^20 (double-buffered inputs and
outputs so that mode 2's promise holds), several rounds."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1] if len(sys.argv) > 1 else "fc"
m = Model.load(os.path.join(root, "tests", "golden", "models", name + ".bnm"))
e = Engine(m)
n = 1 << 20
x = [torch.randint(-128, 128, (n, 256), dtype=torch.int8, device="cuda") for _ in range(2)]
lo = [torch.empty((n, e.n_classes), dtype=torch.int32, device="cuda") for _ in range(2)]
la = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
res = {0: [], 1: [], 2: []}
for rnd in range(5):
    for mode in (0, 1, 2):
        e.set_option(_lib.OPT_LAUNCH_OVERLAP, mode)
        for i in range(5):
            e.infer_device(x[i & 1], lo[i & 1], la[i & 1])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            e.infer_device(x[i & 1], lo[i & 1], la[i & 1])
        b.record(); torch.cuda.synchronize()
        res[mode].append(a.elapsed_time(b) / 40 * 1e3)
for mode in (0, 1, 2):
    print(f"{name} mode {mode}: " + " ".join(f"{v:.2f}" for v in res[mode]) + f"  us/step (median {sorted(res[mode])[2]:.2f})")
# ordinary CUDA semantics, two streams: consecutive (independent, double-buffered) batches alternate between two streams
e.set_option(_lib.OPT_LAUNCH_OVERLAP, 0)
s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
vals = []
for rnd in range(5):
    torch.cuda.synchronize()
    a, b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s2[0]); s2[1].wait_event(a)
    for i in range(40):
        e.infer_device(x[i & 1], lo[i & 1], la[i & 1], s2[i & 1].cuda_stream)
    b0.record(s2[0]); b1.record(s2[1]); torch.cuda.synchronize()
    vals.append(max(a.elapsed_time(b0), a.elapsed_time(b1)) / 40 * 1e3)
print(f"{name} mode 0 on two alternating streams: " + " ".join(f"{v:.2f}" for v in vals) + " us/step")
