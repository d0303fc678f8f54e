import sys, numpy as np
sys.path.insert(0, '.')
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
from oracle.oracle import Oracle
m = Model.load('tests/golden/models/cnn.bnm')
imgs = np.random.default_rng(0).integers(-128, 128, size=(40, 256), dtype=np.int8)
e = Engine(m)
try:
    lo, la = e.infer(imgs)
    wo, wl = Oracle().infer(m, imgs)
    print('match', np.array_equal(lo, wo), np.array_equal(la, wl), (lo != wo).sum(), lo[0], wo[0])
except Exception as ex:
    print('EXC', ex)
