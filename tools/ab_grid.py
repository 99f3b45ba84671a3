"""A/B on one box: cfg-2 training step (graph replay, L2 flushed) with the direction / sine-weight arrays read from HBM
vs computed in the kernels (RENI_FLAG_GRID_DIRECTIONS / RENI_FLAG_GRID_SINEWEIGHT), unmasked and with a rectangle mask."""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, pack_mask_bits, rectangle_mask
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = 32, 36, 128
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
mask = rectangle_mask(W, 10, 46, 40, 82).to(dev)
bits = pack_mask_bits(mask)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = m.Z.detach()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def make(grid, masked):
    ws = F_.Workspace()
    def step():
        ws.prepared_key = None
        if grid:
            return F_.loss_forward_backward(m.spec, ws, Z, None, tg, None, m.decoder_weights(), m.decoder_biases(),
                                            mask_bits=bits if masked else None)
        s = sw * mask if masked else sw  # the per-step torch multiply of RENI_module.py:92-94
        return F_.loss_forward_backward(m.spec, ws, Z, D, tg, s, m.decoder_weights(), m.decoder_biases())
    for _ in range(3): step()
    torch.cuda.synchronize()
    s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_): step()
    torch.cuda.current_stream().wait_stream(s_)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    return g
def timeit(g, n=40):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)
gs = {(g, k): make(g, k) for g in (False, True) for k in (False, True)}
for rep in range(2):
    for (g, k), gr in gs.items():
        print(f"{'analytic grid' if g else 'arrays from HBM'}{', masked' if k else ''}: {timeit(gr)*1e3:.1f} us")
