"""In-graph (warm, replayed) time of the FiLM per-map stage at the cfg-2 shape (32 maps): forward alone, forward +
backward, against the whole training step.  Tells what the per-map stage really costs inside a replayed step (ncu's
per-launch times are cold-cache and serialised)."""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from reni_b200 import RENIAutoDecoderFiLM
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N = int(os.environ.get("RENI_B", "32")), 36
m = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
Z = m.Z
def fwd():
    return m.map_level(Z)
def fwd_bwd():
    mc, film = m.map_level(Z)
    (mc.sum() + film.sum()).backward()
def graph_of(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g
def timeit(g, n=50):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e3
with torch.no_grad():
    g1 = graph_of(fwd)
print(f"per-map stage forward (graph replay, warm): {timeit(g1):.1f} us")
g2 = graph_of(fwd_bwd)
print(f"per-map stage forward + backward (incl. two torch sum kernels + their backward): {timeit(g2):.1f} us")

# per-stage stamps of the fused forward (library built with -DRENI_MAP_FUSED_STAMPS=1, RENI_B200_LIB=...): direct ABI call
if os.environ.get("STAMPS"):
    import ctypes as C
    from reni_b200 import _lib
    from torch import nn
    lib = _lib.load()
    lins = [x for x in m.mapping_network.network if isinstance(x, nn.Linear)]
    ws = [x.weight.detach().float().contiguous() for x in lins]
    bs = [x.bias.detach().float().contiguous() for x in lins]
    dims = (C.c_int32 * (len(lins) + 1))(lins[0].in_features, *[x.out_features for x in lins])
    W0 = m.net[0].layer.weight.detach().float().contiguous(); b0 = m.net[0].layer.bias.detach().float().contiguous()
    mc = torch.empty(B, 5, 256, device=dev); film = torch.empty(B, 4, 2, 256, device=dev)
    nbytes = int(lib.reni_film_map_scratch_bytes(dims, len(lins), B))
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    Zc = Z.detach().float().contiguous()
    for _ in range(5):
        rc = lib.reni_film_map_forward(C.byref(m.spec.c_config()), C.c_void_p(Zc.data_ptr()), C.c_void_p(W0.data_ptr()),
                                       C.c_void_p(b0.data_ptr()), ptrs(ws), ptrs(bs), dims, len(lins), B,
                                       C.c_void_p(mc.data_ptr()), C.c_void_p(film.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                       nbytes, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
    st = scratch[-256:].view(torch.int64).cpu().numpy()
    t = st[2:2 + len(lins) + 3]
    print("fused forward stages (us): input, linears..., finish:", [round((int(b) - int(a)) / 1e3, 2) for a, b in zip(t[:-1], t[1:])])
