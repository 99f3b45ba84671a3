"""N>1 host logic on CPU: world_size-2 gloo processes exercise the flat weight-gradient buffer's
all-reduce (DDP averaging semantics, run.py:97) and the map sharding."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reni_b200 import FlatGradBuffer, shard_range

        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(4)),
                  torch.nn.Parameter(torch.zeros(2, 4))]
        fb = FlatGradBuffer(params)
        assert fb.flat.numel() == 12 + 4 + 8
        # per-rank "batch-summed" gradients written through the views (what the kernels accumulate into)
        for i, v in enumerate(fb.views):
            v += float(rank + 1) * (i + 1)
        fb.all_reduce_mean()
        fb.attach()
        expect = [sum(r + 1 for r in range(world)) / world * (i + 1) for i in range(3)]
        ok = all(torch.allclose(p.grad, torch.full_like(p, e)) for p, e in zip(params, expect))
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, fb.views))
        lo, hi = shard_range(33, rank, world)
        sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        ok = ok and int(sum(s.item() for s in sizes)) == 33
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _latent_worker(rank, world, port, q):
    """Host logic of the latent tables under data parallelism (no kernels run): shard check, replicated gradient sync
    (reference DDP: the dense table gradient is averaged over ranks) and gather_latents()."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reni_b200 import RENIAutoDecoder, RENITrainer, shard_range

        torch.manual_seed(0)
        m = RENIAutoDecoder(8, 4, "SO2", 256, 1, 3, True, "tanh", 30.0, 30.0, False)  # CPU: only host logic is used
        ok = True
        tr = RENITrainer(m, "FIT_DECODER", 16, lr=1e-3)  # latent_sync="local"
        lo, hi = shard_range(8, rank, world)
        tr._check_shard(torch.arange(lo, hi))            # own rows: fine
        try:
            tr._check_shard(torch.tensor([(hi) % 8]))    # a row of the other rank
            ok = False
        except ValueError:
            pass
        # gather_latents: every rank changes only its own rows; the gathered table holds everybody's
        with torch.no_grad():
            m.Z[lo:hi] = float(rank + 1)
        full = tr.gather_latents()["Z"]
        for r in range(world):
            a, b = shard_range(8, r, world)
            ok = ok and bool((full[a:b] == float(r + 1)).all())
        # replicated: local gradient (already divided by the world size) summed over ranks == DDP's average
        tr2 = RENITrainer(m, "FIT_DECODER", 16, lr=1e-3, latent_sync="replicated")
        tr2._check_shard(torch.tensor([(hi) % 8]))       # any row is allowed
        g = torch.zeros_like(m.Z)
        g[lo:hi] = float(rank + 1) / world
        m.Z.grad = g
        tr2._sync_latent_grads()
        for r in range(world):
            a, b = shard_range(8, r, world)
            ok = ok and torch.allclose(m.Z.grad[a:b], torch.full_like(m.Z.grad[a:b], float(r + 1) / world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_latent_tables_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_latent_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
