"""Parity of the CNN path against the oracle (run with BNM_CNN_TC=1 to exercise cnn_tcgen05.cu); prints one line."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
from oracle.oracle import Oracle
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ok_l = ok_b = True; bad = 0; first = None
for name, n in (("cnn", 1027), ("cnn_48", 643)):
    m = Model.load(os.path.join(root, "tests", "golden", "models", name + ".bnm"))
    imgs = np.random.default_rng(len(name)).integers(-128, 128, size=(n, 256), dtype=np.int8)
    e = Engine(m)
    lo, la = e.infer(imgs)
    wo, wl = Oracle().infer(m, imgs)
    e.close()
    ok_l &= bool(np.array_equal(lo, wo)); ok_b &= bool(np.array_equal(la, wl)); bad += int((lo != wo).sum())
    first = first if first is not None else lo[0].tolist()
print("match", ok_l, ok_b, bad, first)
