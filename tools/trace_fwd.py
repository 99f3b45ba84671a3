"""Pipeline timeline of CTA 0 of the training forward kernel (BASELINE configs[1]): clock64 stamps from the MMA issuer
and one warp of each epilogue group (library debug hook reni_debug_set_trace)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
    ws.prepared_key = None
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
names = {9: "clk_mark", 10: "wall_ns", 1: "A_seen", 2: "issued", 3: "wait_acc", 4: "acc_seen", 5: "epi_done", 6: "w_wait", 7: "w_own", 8: "w_all"}
rows, walls = [], []
for r in range(3):
    for x in ev[r]:
        if x == 0: continue
        code, clk = int(x >> np.uint64(48)), int(x & np.uint64(0xFFFFFFFFFFFF))
        if (code >> 8) == 10: walls.append(clk)
        else: rows.append((clk, r, names.get(code >> 8, "?"), (code >> 4) & 15, code & 15))
rows.sort()
t0 = rows[0][0]
lo, hi = int(os.environ.get("LO", "200")), int(os.environ.get("HI", "420"))
prev = {}
for clk, r, nm, l, g in rows[lo:hi]:
    role = ["mma ", "epi0", "epi1"][r]
    print(f"{clk - t0:9d}  {role} {nm:9s} l={l} g={g}")

ck = sorted(c for c, r, nm, l, g in rows if nm == "clk_mark"); wl = sorted(walls)
if len(ck) >= 2 and len(wl) >= 2:
    print(f"issuer span: {ck[-1] - ck[0]} clk in {(wl[-1] - wl[0]) / 1e3:.1f} us -> SM clock {(ck[-1] - ck[0]) / (wl[-1] - wl[0]) * 1e3:.0f} MHz")
