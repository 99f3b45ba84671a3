"""Latency of the gradient exchange alone (torchrun, one rank per GPU): K back-to-back all-reduces of the training-size
flat buffer, amortising the launch skew between ranks.  reni_allreduce (multicast / p2p, captured in one CUDA graph) vs
ncclAllReduce(avg) issued back to back.   python -m torch.distributed.run --nproc-per-node N tools/exchange_bench.py [numel]"""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from reni_b200.training import FlatGradBuffer
numel = int(sys.argv[1]) if len(sys.argv) > 1 else 680707
K = 50
params = [torch.nn.Parameter(torch.zeros(numel, device=dev))]
res = {}
for mode in ("multicast", "p2p", "nccl"):
    os.environ["RENI_EXCHANGE"] = mode
    try:
        fb = FlatGradBuffer(params)
    except Exception as e:
        res[mode] = f"unavailable: {e!r}"
        continue
    fb.flat.fill_(1.0)
    for _ in range(3):
        fb.all_reduce_mean()
    torch.cuda.synchronize()
    if fb.capturable:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fb.all_reduce_mean()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(K):
                fb.all_reduce_mean()
        run = g.replay
    else:
        def run():
            for _ in range(K):
                fb.all_reduce_mean()
    ts = []
    for _ in range(5):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / K * 1e3)
    ok = bool(torch.allclose(fb.flat, torch.ones_like(fb.flat))) and not fb.exchange_failed()
    res[mode] = f"{min(ts):.1f} us per all-reduce (best of 5 x {K}), exchange={fb.exchange}, result ok={ok}"
if rank == 0:
    print(f"world {world}, {numel} fp32 ({numel*4/1e6:.2f} MB):")
    for k, v in res.items():
        print("  ", k, v)
dist.destroy_process_group()
