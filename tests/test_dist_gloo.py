"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path (column sharding, gradient averaging,
max-over-ranks timing reduction) -- the N>1 plumbing of bench.py without GPUs."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from regneuralde.jl_b200.parallel import shard_columns, average_gradients_, max_over_ranks
    B = 10
    lo, hi = shard_columns(B, rank, world)
    # every column owned exactly once
    owned = torch.zeros(B); owned[lo:hi] = 1
    dist.all_reduce(owned)
    # gradient averaging == mean of the per-rank gradients (loss is a mean over the global batch)
    g = torch.full((5,), float(rank + 1))
    average_gradients_([g], world)
    t = max_over_ranks(float(rank + 3))
    # rendezvous of the reference-exact mode: every rank contributes the 64-byte IPC handle of its exchange buffer and
    # receives all of them in rank order (rnde_dist_export -> exchange_ipc_handles -> rnde_dist_import)
    from regneuralde.jl_b200.parallel import exchange_ipc_handles
    hs = exchange_ipc_handles(bytes([rank + 1]) * 64, world)
    assert [h[0] for h in hs] == list(range(1, world + 1)) and all(len(h) == 64 for h in hs)
    q.put((rank, lo, hi, owned.tolist(), g.tolist(), t))
    dist.destroy_process_group()


def test_sharding_and_gradient_average_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, own0, g0, t0), (r1, lo1, hi1, own1, g1, t1) = out
    assert (lo0, hi0, lo1, hi1) == (0, 5, 5, 10)
    assert own0 == [1.0] * 10 and own1 == [1.0] * 10
    assert g0 == [1.5] * 5 and g1 == [1.5] * 5          # (1 + 2) / 2
    assert t0 == 4.0 and t1 == 4.0                       # max over ranks


def test_shard_columns_ragged():
    from regneuralde.jl_b200.parallel import shard_columns
    parts = [shard_columns(513, r, 8) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 513
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
