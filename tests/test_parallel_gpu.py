"""Data-parallel numerics on real GPUs (needs >= 2 devices; skipped otherwise): two NCCL ranks, one process per GPU.

  * reni_allreduce (the in-graph exchange over symmetric memory: two-shot over peer pointers, and multimem on the
    multicast address where the fabric has one) == NCCL all-reduce(avg) of the same buffers;
  * a 2-rank RENITrainer step == the 1-rank step on the concatenated batch: weight gradients are the average of the
    per-rank batch sums (reference DDP, run.py:97), latent gradients are divided by the world size, and the losses of
    the two shards add up to the single-rank loss; eager and CUDA-graph replay (exchange captured in the graph)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT

pytestmark = pytest.mark.gpu


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
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    msgs = []
    ok = True
    try:
        from reni_b200 import FlatGradBuffer, RENIAutoDecoder, RENITrainer, shard_range

        # ---- (1) exchange kernel vs NCCL on a buffer of the training size
        torch.manual_seed(100 + rank)
        params = [torch.nn.Parameter(torch.zeros(256, 1370, device=dev)), torch.nn.Parameter(torch.zeros(256, device=dev)),
                  torch.nn.Parameter(torch.zeros(5, 256, 256, device=dev)), torch.nn.Parameter(torch.zeros(771, device=dev))]
        for mode in ("p2p", "multicast"):
            os.environ["RENI_EXCHANGE"] = mode
            try:
                fb = FlatGradBuffer(params)
            except Exception as e:  # no multicast on this fabric: the p2p path must still exist
                msgs.append(f"{mode}: unavailable ({e!r})")
                ok = ok and mode == "multicast"
                continue
            finally:
                os.environ.pop("RENI_EXCHANGE", None)
            for rep in range(3):
                fb.flat.copy_(torch.randn(fb.flat.numel(), device=dev) * (rep + 1))
                ref = fb.flat.clone()
                dist.all_reduce(ref, op=dist.ReduceOp.AVG)
                fb.all_reduce_mean()
                torch.cuda.synchronize()
                err = float((fb.flat - ref).abs().max())
                ok = ok and err <= 1e-6 * float(ref.abs().max()) and not fb.exchange_failed()
            msgs.append(f"{mode}: exchange={fb.exchange} ok")

        # ---- (2) 2-rank trainer step vs the 1-rank step on the concatenated batch
        for graph in (False, True):
            torch.manual_seed(0)
            W, B = 32, 4  # 4 maps per rank
            total = B * world
            m = RENIAutoDecoder(total, 9, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
            g = torch.Generator().manual_seed(5)
            imgs_all = (torch.rand(total, 3, W // 2, W, generator=g) * 2 - 1).to(dev)
            lo, hi = shard_range(total, rank, world)
            tr = RENITrainer(m, "FIT_DECODER", W, lr=1e-4, cuda_graph=graph)
            for _ in range(2):
                log = tr.training_step((imgs_all[lo:hi], torch.arange(lo, hi)))
            torch.cuda.synchronize()
            dw = [p.grad.clone() for p in m.net.parameters()]
            dz = m.Z.grad.clone()
            loss = log["loss"].clone().reshape(1)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
            # single-rank reference on this rank's GPU (no process group inside: world_size forced to 1)
            m1 = RENIAutoDecoder(total, 9, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
            m1.load_state_dict({"model." + k: v for k, v in m.state_dict().items()})  # (Lightning-prefixed keys, RENI.py:190-203)
            tr1 = RENITrainer(m1, "FIT_DECODER", W, lr=1e-4)
            tr1.world_size = 1
            tr1.flat.exchange = "none"
            tr1.flat._symm = None
            log1 = tr1.training_step((imgs_all, torch.arange(total)))
            torch.cuda.synchronize()
            ok = ok and abs(float(loss) - float(log1["loss"])) <= 1e-5 * abs(float(log1["loss"]))
            for a, p in zip(dw, m1.net.parameters()):
                e = float((a - p.grad / world).norm() / (p.grad / world).norm())
                ok = ok and e < 2e-4
            ez = float((dz[lo:hi] - m1.Z.grad[lo:hi] / world).norm() / (m1.Z.grad[lo:hi] / world).norm())
            ok = ok and ez < 2e-4 and float(dz[:lo].abs().sum() + dz[hi:].abs().sum()) == 0.0
            msgs.append(f"trainer graph={graph}: exchange={tr.flat.exchange} in_graph={tr.flat.capturable} ok={ok}")
        q.put((rank, bool(ok), msgs))
    except Exception as e:  # surface the traceback to the parent
        import traceback

        q.put((rank, False, msgs + [traceback.format_exc()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_exchange_and_trainer_match_single_rank():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import __graft_entry__ as entry

    entry.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msgs in sorted(res):
        print(f"rank {rank}:", *msgs, sep="\n   ")
    assert all(ok for _, ok, _ in res), res
