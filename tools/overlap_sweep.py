"""Overlap mode of the training backward (weight-gradient kernel co-resident with the delta chain): gradients against
the back-to-back kernels, then step time (CUDA-graph replay, L2 flushed between steps) per SM share given to dW."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, _lib
from reni_b200 import functional as F_
from reni_b200.training import FlatGradBuffer
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = int(os.environ.get("B", 32)), int(os.environ.get("N", 36)), int(os.environ.get("W", 128))
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = torch.randn(B, N, 3, device=dev)
ws = F_.Workspace()
flat = FlatGradBuffer(m.decoder_parameters())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
def step():
    ws.prepared_key = None
    flat.zero_()
    r = F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), need_dw=True,
                                 grad_weights=flat.views[0::2], grad_biases=flat.views[1::2])
    res["r"] = r
def timeit(fn, n=40):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) * 1e3, min(ts) * 1e3
lib = _lib.load()
out = {}
ref = None
settings = [(0, 0)] + [(int(a), int(b)) for a, b in (s.split(":") for s in os.environ.get(
    "SWEEP", "-1:0,40:0,48:0,52:0,56:0,60:0,66:0,52:4,56:4").split(","))]
for dw, oc in settings:
    assert lib.reni_debug_set_overlap(dw, oc) == 0
    for _ in range(3): step()
    torch.cuda.synchronize()
    gflat = flat.flat.clone()
    if ref is None:
        ref = gflat
        err = 0.0
    else:
        err = float((gflat - ref).norm() / ref.norm())
    eager = timeit(step, 20)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    gerr = float((flat.flat - ref).norm() / ref.norm())
    med, best = timeit(g.replay)
    out[f"{dw}:{oc}"] = {"graph_us": round(med, 1), "graph_best_us": round(best, 1), "eager_us": round(eager[0], 1),
                         "dW_rel_vs_sequential": err, "dW_rel_graph": gerr}
    print(dw, oc, out[f"{dw}:{oc}"], flush=True)
lib.reni_debug_set_overlap(-1, 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"B": B, "N": N, "W": W, "results": out}, open(os.path.join(ROOT, "gpurun_out", "overlap_sweep.json"), "w"), indent=1)
