"""MMA-issuer timeline of CTA 0 of the training forward (cfg 2): per (layer, sub-tile) pass, how long the issuer waited
for the epilogue's operand (a_ready) and how long the pass took to issue (weight ring waits included)."""
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
lib = _lib.load()
for terms in ("1", "2"):
    os.environ["RENI_FWD_TERMS"] = terms
    ws = F_.Workspace()
    def step():
        ws.prepared_key = None
        F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), need_dw=True, tile_major_bwd=True)
    for _ in range(3): step()
    torch.cuda.synchronize()
    buf = torch.zeros(3 * 4096, dtype=torch.int64, device=dev)
    lib.reni_debug_set_trace(C.c_void_p(buf.data_ptr()))
    step(); torch.cuda.synchronize()
    lib.reni_debug_set_trace(None)
    ev = buf.cpu().numpy().astype(np.uint64).reshape(3, 4096)[0]
    rows = [(int(x >> np.uint64(48)), int(x & np.uint64(0xFFFFFFFFFFFF))) for x in ev if x]
    # code: 0x100|(l<<4)|g = operand ready seen ; 0x200|(l<<4)|g = all MMAs of the pass issued
    waits, issues = [], []
    prev_issued = None
    for code, clk in rows:
        kind = code >> 8
        if kind == 1:
            if prev_issued is not None: waits.append(clk - prev_issued)
            t_seen = clk
        elif kind == 2:
            issues.append((clk - t_seen, (code >> 4) & 15))
            prev_issued = clk
    hid = [d for d, l in issues if 1 <= l <= 5]
    print(f"terms={terms}: passes {len(issues)}, hidden-layer pass issue time median {np.median(hid):.0f} clk (min {min(hid)}, max {max(hid)}); "
          f"wait for operand median {np.median(waits):.0f} clk; total span {rows[-1][1]-rows[0][1]} clk")
    seq = [f"{'L'+str((c>>4)&15)+'g'+str(c&15)}:{'seen' if c>>8==1 else 'iss'}@{clk-rows[0][1]}" for c, clk in rows[40:76]]
    print("   ", " ".join(seq))
