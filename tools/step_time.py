"""Whole-step device time of BASELINE configs[1], three ways: eager launches (fork/join active), CUDA-graph replay,
and graph replay without the L2 flush between steps."""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight
from reni_b200 import functional as F_
from reni_b200.training import FlatGradBuffer
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = 32, 36, 128
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = torch.randn(B, N, 3, device=dev)
ws = F_.Workspace()
flat = FlatGradBuffer(m.decoder_parameters())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    ws.prepared_key = None
    flat.zero_()
    F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), need_dw=True,
                             grad_weights=flat.views[0::2], grad_biases=flat.views[1::2])
def timeit(fn, n=60, do_flush=True):
    ts = []
    for _ in range(n):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts) * 1e3
for _ in range(5): step()
torch.cuda.synchronize()
print(f"eager (fork/join active)      {timeit(step):7.0f} us")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
for _ in range(3): g.replay()
print(f"graph replay                  {timeit(g.replay):7.0f} us")
print(f"graph replay, no L2 flush     {timeit(g.replay, do_flush=False):7.0f} us")
