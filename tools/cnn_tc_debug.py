"""CNN front-end check in one process per kernel: parity against the oracle on ragged batches, then device-timed throughput.

    python tools/cnn_tc_debug.py <frontend: 1 CUDA cores | 2 tensor cores> [log2 batch for the timing, default 18]
"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
from oracle.oracle import Oracle
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fe = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 18
orc = Oracle()
for name, n in (("cnn", 1027), ("cnn_48", 643), ("cnn_32", 517), ("cnn_16", 1001)):
    m = Model.load(os.path.join(root, "tests", "golden", "models", name + ".bnm"))
    imgs = np.random.default_rng(len(name)).integers(-128, 128, size=(n, 256), dtype=np.int8)
    e = Engine(m)
    e.set_option(_lib.OPT_CNN_FRONTEND, fe)
    lo, la = e.infer(imgs)
    wo, wl = orc.infer(m, imgs)
    bad = np.argwhere(lo != wo)
    print(f"frontend {fe} {name}: logits {bool(np.array_equal(lo, wo))} labels {bool(np.array_equal(la, wl))} mismatching {len(bad)} first {bad[:3].tolist()}", flush=True)
    e.close()
for name in ("cnn", "cnn_48"):
    m = Model.load(os.path.join(root, "tests", "golden", "models", name + ".bnm"))
    n = 1 << lg
    e = Engine(m)
    e.set_option(_lib.OPT_CNN_FRONTEND, fe)
    d_in = torch.randint(-128, 128, (n, 256), dtype=torch.int8, device="cuda")
    d_log = torch.empty((n, e.n_classes), dtype=torch.int32, device="cuda")
    d_lab = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        e.infer_device(d_in, d_log, d_lab)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        e.infer_device(d_in, d_log, d_lab)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"frontend {fe} {name}: batch 2^{lg} {ms:.3f} ms/step = {n / ms / 1e6:.4f} G images/s", flush=True)
    e.close()
