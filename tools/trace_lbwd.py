"""Pipeline timeline of CTA 0 of one layer-major backward launch (layer L-1) at BASELINE configs[1]: clock64 stamps from
the MMA issuer, the first epilogue warp and the producer (library debug hook reni_debug_set_trace)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, _lib
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, W = 32, 36, 128
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = torch.randn(B, N, 3, device=dev)
ws = F_.Workspace()
def step():
    F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), need_dw=True)
for _ in range(3): step()
torch.cuda.synchronize()
buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.reni_debug_set_trace(C.c_void_p(buf.data_ptr()))
step()
torch.cuda.synchronize()
lib.reni_debug_set_trace(None)
ev = buf.cpu().numpy().astype(np.uint64).reshape(3, 4096)
names = {0: {1: "tile_landed", 2: "chain_issue", 3: "wgrad_issue"},
         1: {1: "iter_start", 2: "h_free_seen", 3: "h_written", 4: "acc_seen", 5: "cos_done"},
         2: {1: "load_issue"}}
rows = []
for r in range(3):
    for x in ev[r]:
        if x == 0: continue
        code, clk = int(x >> np.uint64(48)), int(x & np.uint64(0xFFFFFFFFFFFF))
        rows.append((clk, r, names[r].get(code >> 8, "?"), code & 255))
rows.sort()
rows = [x for x in rows if x[2] != "?"]
t0 = rows[0][0]
for clk, r, nm, i in rows[: int(os.environ.get("HI", "400"))]:
    print(f"{clk - t0:9d}  {['mma ', 'epi ', 'prod'][r]} {nm:12s} tile {i}")
