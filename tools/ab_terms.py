"""A/B on one box: cfg-2 training step (CUDA-graph replay, L2 flushed) with one-term vs two-term forward weights and
tile-major vs layer-major backward.  Prints ms per step for each combination, interleaved twice."""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = 32, 36, 128
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = m.Z.detach()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def make(terms, tile_major):
    ws = F_.Workspace()
    os.environ["RENI_FWD_TERMS"] = str(terms)
    def step():
        ws.prepared_key = None
        return F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), tile_major_bwd=tile_major)
    for _ in range(3): step()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): r = step()
    os.environ.pop("RENI_FWD_TERMS")
    return g, r
def timeit(g, n=30):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
combos = [(1, True), (2, True), (1, False), (2, False)]
graphs = {c: make(*c) for c in combos}
ref = graphs[(2, True)][1].out.clone()
for c in combos:
    graphs[c][0].replay(); torch.cuda.synchronize()
    print(c, "out differs from two-term:", float((graphs[c][1].out - ref).abs().max()))
for rep in range(2):
    for c in combos:
        print(f"terms={c[0]} {'tile-major' if c[1] else 'layer-major'}: {timeit(graphs[c][0])*1e3:.1f} us")
# sustained regime (what bench.py reports as sustained_ms_per_step): 300 back-to-back replays, mean of the last 100 --
# the boxes run against a power cap, so traffic saved may count for more here than in isolated steps
def sustained(g, n=300):
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for i in range(n):
        flush.zero_()
        e0[i].record(); g.replay(); e1[i].record()
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in zip(e0, e1)]
    return sum(ts[-100:]) / 100
if os.environ.get("SUSTAIN", "1") != "0":
    for rep in range(2):
        for c in combos:
            print(f"sustained terms={c[0]} {'tile-major' if c[1] else 'layer-major'}: {sustained(graphs[c][0])*1e3:.1f} us")
