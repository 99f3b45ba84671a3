"""Latency of decoding a few latents into 64x128 environment maps (no grad): op-by-op eager call vs GraphedDecoder
(one CUDA-graph replay).  Wall-clock per call including the host side, median of 200 calls, each followed by a sync."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from reni_b200 import GraphedDecoder, RENIAutoDecoder, RENIAutoDecoderFiLM, get_directions
dev = torch.device("cuda:0")
torch.manual_seed(0)
W = 128
D = get_directions(W).to(dev)
out = {}
def wall(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6
for name, mk in (("concat", lambda B: RENIAutoDecoder(B, 36, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, True)),
                 ("film", lambda B: RENIAutoDecoderFiLM(B, 36, "SO2", 256, 5, 256, 3, 3, None, True))):
    for B in (1, 8, 32):
        m = mk(B).to(dev)
        Z = torch.randn(B, 36, 3, device=dev)
        Db = D.expand(B, -1, -1)
        def eager():
            with torch.no_grad(): return m(Z, Db)
        gd = GraphedDecoder(m, B, D)
        out[f"{name}_B{B}"] = {"eager_us": round(wall(eager), 1), "graph_replay_us": round(wall(lambda: gd(Z)), 1)}
        print(name, B, out[f"{name}_B{B}"], flush=True)
print(json.dumps(out, indent=1))
