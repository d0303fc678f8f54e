"""Cost of the fused-gather kernel variant with destinations on the SAME GPU (no NVLink): us per step at batch 2^20."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
e = Engine(Model.load(os.path.join(root, "tests", "golden", "models", "fc.bnm")))
n = 1 << 20
x = [torch.randint(-128, 128, (n, 256), dtype=torch.int8, device="cuda") for _ in range(2)]
lo = [torch.empty((n, 10), dtype=torch.int32, device="cuda") for _ in range(2)]
la = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(2)]
glab = [torch.empty(n * 8, dtype=torch.int32, device="cuda") for _ in range(7)]
glog = [torch.empty((n * 2, 10), dtype=torch.int32, device="cuda") for _ in range(2)]
st = torch.cuda.current_stream().cuda_stream
def run(n_lab, n_log, u8=0):
    g = _lib.BnmGather(); g.n_labels_dst = n_lab; g.n_logits_dst = n_log; g.row_offset = 0; g.labels_u8 = u8
    for k in range(n_lab): g.labels_dst[k] = glab[k].data_ptr()
    for k in range(n_log): g.logits_dst[k] = glog[k].data_ptr()
    def step(i):
        if n_lab or n_log:
            _lib.check(e.lib.bnm_infer_batch_device_gather(e.handle, C.c_void_p(x[i & 1].data_ptr()), n, C.c_void_p(lo[i & 1].data_ptr()), C.c_void_p(la[i & 1].data_ptr()), C.byref(g), C.c_void_p(st)), "g")
        else:
            e.infer_device(x[i & 1], lo[i & 1], la[i & 1])
    for i in range(5): step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40): step(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 40 * 1e3
for rnd in range(2):
    print("plain %.2f | 1 label dst %.2f | 7 label dst %.2f | 7 label dst u8 %.2f | 1 logits dst %.2f | 2 logits dst %.2f  us/step" %
          (run(0, 0), run(1, 0), run(7, 0), run(7, 0, 1), run(0, 1), run(0, 2)))
