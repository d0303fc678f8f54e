"""The multi-process half of the fused result exchange (bitnetmcu_b200/gather.py): two ranks, gather buffers exchanged as CUDA IPC
handles over torch.distributed (gloo), each rank's kernel storing its rows into the other rank's buffers.  Both processes share
GPU 0 here -- CUDA IPC works between processes on one device -- so the path is exercised on a single-GPU box; with two GPUs the
ranks take one each (then the stores really cross NVLink)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["BNM_ROOT"])
from bitnetmcu_b200 import _lib
from bitnetmcu_b200.engine import Engine
from bitnetmcu_b200.gather import PeerGatherBuffers, infer_gather
from bitnetmcu_b200.model import Model
from oracle.oracle import Oracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
dev = rank if torch.cuda.device_count() >= world else 0
torch.cuda.set_device(dev)
m = Model.load(os.path.join(os.environ["BNM_ROOT"], "tests", "golden", "models", "fc.bnm"))
n = 148 * 128 + 333                                   # full tiles (staged bulk stores) + a ragged tail (direct stores)
imgs = [np.random.default_rng(100 + r).integers(-128, 128, size=(n, 256)).astype(np.int8) for r in range(world)]
want = [Oracle().infer(m, x) for x in imgs]
e = Engine(m, device=dev)
bufs = PeerGatherBuffers(n, 10, dev, with_logits=True)
bufs.logits.fill_(-1); bufs.labels.fill_(-1); bufs.labels_u8.fill_(255)
torch.cuda.synchronize(); dist.barrier()
d_img = torch.from_numpy(imgs[rank]).to(f"cuda:{dev}")
peers = [r for r in range(world) if r != rank]
st = torch.cuda.current_stream().cuda_stream
my_log, my_lab = bufs.logits[rank * n:(rank + 1) * n], bufs.labels[rank * n:(rank + 1) * n]
infer_gather(e, d_img, my_log, my_lab, bufs, peers, peers, st)                                  # uint32 labels + logits to the peers
infer_gather(e, d_img, my_log, my_lab, bufs, list(range(world)), None, st, labels_u8=True)      # one-byte labels to everybody
torch.cuda.synchronize(); dist.barrier()
ok = True
for r in range(world):
    ok &= bool(np.array_equal(bufs.logits[r * n:(r + 1) * n].cpu().numpy(), want[r][0]))
    ok &= bool(np.array_equal(bufs.labels[r * n:(r + 1) * n].cpu().numpy().astype(np.uint32), want[r][1]))
    ok &= bool(np.array_equal(bufs.labels_u8[r * n:(r + 1) * n].cpu().numpy(), want[r][1].astype(np.uint8)))
dist.barrier()
bufs.close(); e.close()
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 3)
'''


def test_two_process_fused_gather_over_cuda_ipc(built, tmp_path):
    from bitnetmcu_b200 import _lib
    assert _lib.load().bnm_device_count() > 0
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", BNM_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    report = "\n".join(f"---- rank {r} (exit {p.returncode})\n{o[-2500:]}" for r, (p, o) in enumerate(zip(procs, outs)))
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK {r} OK" in o, report
