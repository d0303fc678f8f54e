#!/usr/bin/env python
"""BASELINE.json config 5: encoding sweep x batch sweep on one GPU (device-timed, CUDA events, inputs resident in HBM).
Prints a markdown table; every point is parity-checked against the oracle on a 4096-image sample (checker only)."""
import os, sys, statistics
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.model import Model
from oracle.oracle import Oracle

orc = Oracle()
PEAK = 6592.9


def run(name, n, steps=10, nf4=False):
    m = Model.load(os.path.join(ROOT, "tests", "golden", "models", name + ".bnm"))
    e = Engine(m, nf4_extension=nf4)
    rng = np.random.default_rng(1)
    h = rng.integers(-128, 128, size=(min(n, 1 << 22), 256), dtype=np.int8)
    x = torch.from_numpy(h).cuda()
    if n > x.shape[0]:
        x = x.repeat(n // x.shape[0], 1)
    lo = torch.empty((n, e.n_classes), dtype=torch.int32, device="cuda")
    la = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        e.infer_device(x, lo, la)
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e.infer_device(x, lo, la); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ns = min(n, 4096)
    want, wl = orc.infer(m, h[:ns], nf4_extension=nf4)
    ok = np.array_equal(lo[:ns].cpu().numpy(), want) and np.array_equal(la[:ns].cpu().numpy().astype(np.uint32), wl)
    ms = statistics.median(ts)
    e.close()
    return n / ms / 1e6, (256 + 4 * m.n_classes) * n / ms / 1e6 / PEAK, ms, ok


print("| model | batch | G images/s | of measured HBM roofline | ms/launch | parity |\n|---|---|---|---|---|---|")
for n in [1 << 10, 1 << 14, 1 << 17, 1 << 20, 1 << 22, 1 << 24]:
    g, f, ms, ok = run("fc", n)
    print(f"| fc (4bitsym-64) | 2^{n.bit_length()-1} | {g:.3f} | {f:.3f} | {ms:.4f} | {'bit-exact' if ok else 'MISMATCH'} |", flush=True)
for name, nf4 in [("rand_binary64", False), ("ternary64", False), ("1k", False), ("2bitsym96", False), ("4bit64", False), ("12k_FP130", False),
                  ("rand_nf4_64", True), ("8bit64", False), ("binary160", False), ("cnn", False), ("cnn_48", False)]:
    g, f, ms, ok = run(name, 1 << 20)
    print(f"| {name}{' (NF4 LUT extension)' if nf4 else ''} | 2^20 | {g:.3f} | {f:.3f} | {ms:.4f} | {'bit-exact' if ok else 'MISMATCH'} |", flush=True)
