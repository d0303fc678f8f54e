#!/bin/bash
# e2e (pinned host buffers through bnm_infer_batch) as a function of the host pipeline's chunk size
mkdir -p gpurun_out
timeout 300 python - <<'PY' | tee gpurun_out/e2e_chunks.log
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.')
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
import torch
m = Model.load('tests/golden/models/fc.bnm'); e = Engine(m); lib = _lib.load()
n = 1 << 20
p_in, p_log, p_lab = lib.bnm_host_alloc(n * 256), lib.bnm_host_alloc(n * 40), lib.bnm_host_alloc(n * 4)
h_in = np.ctypeslib.as_array((C.c_int8 * (n * 256)).from_address(p_in)).reshape(n, 256)
h_log = np.ctypeslib.as_array((C.c_int32 * (n * 10)).from_address(p_log)).reshape(n, 10)
h_lab = np.ctypeslib.as_array((C.c_uint32 * n).from_address(p_lab))
h_in[:] = np.random.default_rng(0).integers(-128, 128, size=(n, 256), dtype=np.int8)
for chunk in [1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20]:
    e.set_option(_lib.OPT_CHUNK_IMAGES, chunk)
    for _ in range(3): e.infer(h_in, out_logits=h_log, out_labels=h_lab)
    t0 = time.perf_counter()
    for _ in range(10): e.infer(h_in, out_logits=h_log, out_labels=h_lab)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"chunk 2^{chunk.bit_length()-1:2d}: {dt*1e3:7.3f} ms per 1M images  {n/dt/1e6:7.1f} M img/s  H2D {n*256/dt/1e9:5.1f} GB/s")
# copies alone, for reference
d = torch.empty(n * 256, dtype=torch.int8, device='cuda'); hp = torch.from_numpy(h_in.reshape(-1))
for _ in range(3): d.copy_(hp, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(hp, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"one 268 MB cudaMemcpyAsync H2D alone: {dt*1e3:.3f} ms = {n*256/dt/1e9:.1f} GB/s")
PY
