import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
import torch
m = Model.load('tests/golden/models/fc.bnm'); e = Engine(m); lib = _lib.load()
N = 1 << 18
p_in, p_log, p_lab = lib.bnm_host_alloc(N * 256), lib.bnm_host_alloc(N * 40), lib.bnm_host_alloc(N * 4)
h_in = np.ctypeslib.as_array((C.c_int8 * (N * 256)).from_address(p_in)).reshape(N, 256)
h_log = np.ctypeslib.as_array((C.c_int32 * (N * 10)).from_address(p_log)).reshape(N, 10)
h_lab = np.ctypeslib.as_array((C.c_uint32 * N).from_address(p_lab))
h_in[:] = 1
def t(n, chunk, labels=True):
    e.set_option(_lib.OPT_CHUNK_IMAGES, chunk)
    for _ in range(2): e.infer(h_in[:n], out_logits=h_log[:n], out_labels=h_lab[:n] if labels else None, want_labels=labels)
    t0 = time.perf_counter()
    for _ in range(5): e.infer(h_in[:n], out_logits=h_log[:n], out_labels=h_lab[:n] if labels else None, want_labels=labels)
    dt = (time.perf_counter() - t0) / 5
    print(f"n {n:7d} chunk {chunk:7d} labels {labels}: {dt*1e3:8.3f} ms  ({dt*1e6/max(1,(n+chunk-1)//chunk):8.1f} us per chunk)")
for n, chunk in [(16384, 65536), (16384, 16384), (18944, 65536), (19072, 65536), (32768, 16384), (65536, 16384), (65536, 18944), (65536, 19072), (65536, 32768), (262144, 16384)]:
    t(n, chunk)
t(65536, 16384, labels=False)
