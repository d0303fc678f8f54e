"""Host-side logic of the N>1 path on CPU: world_size-2 gloo process groups exercise the shard arithmetic and the
gather plumbing of bitnetmcu_b200.dist (the per-rank compute is a stand-in: the engine itself has no CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from bitnetmcu_b200.dist import shard_range
    for n in (0, 1, 7, 128, 1000, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _fake_infer(images):
    """stand-in for Engine.infer: deterministic per-image 'logits' so that order mistakes are visible"""
    x = images.astype(np.int64)
    logits = np.stack([x[:, :16].sum(1) * (k + 1) + k for k in range(10)], axis=1).astype(np.int32)
    return logits, logits.argmax(1).astype(np.uint32)


def _worker(rank, world, port, n, mode, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from bitnetmcu_b200 import dist as bdist
    bdist.init_process_group("gloo")
    rng = np.random.default_rng(0)
    images = rng.integers(-128, 128, size=(n, 256)).astype(np.int8)
    logits, labels = bdist.sharded_infer(_fake_infer, images, gather=mode)
    want_logits, want_labels = _fake_infer(images)
    b, e = bdist.shard_range(n, rank, world)
    if mode == "logits":
        ok = np.array_equal(logits, want_logits) and np.array_equal(labels, want_labels)
    elif mode == "labels":
        ok = np.array_equal(logits, want_logits[b:e]) and np.array_equal(labels, want_labels)
    else:
        ok = np.array_equal(logits, want_logits[b:e]) and np.array_equal(labels, want_labels[b:e])
    q.put((rank, bool(ok)))
    tdist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,mode", [(1000, "logits"), (1001, "logits"), (257, "labels"), (64, "none"), (1, "logits")])
def test_sharded_infer_world2_gloo(n, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
