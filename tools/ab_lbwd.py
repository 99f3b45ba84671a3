"""A/B of the two backward schedules: relative differences of every gradient, several repetitions (debug tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import RENIAutoDecoder
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
def rel(a, b): return float((a - b).norm() / b.norm())
for (B, P, N, L, last_lin, cosine) in [(3, 8192, 9, 5, True, False), (2, 1000, 36, 5, True, True), (32, 8192, 36, 5, True, False)]:
    torch.manual_seed(11)
    m = RENIAutoDecoder(B, N, "SO2", 256, L, 3, last_lin, "tanh", 30.0, 30.0, False).to(dev)
    rng = np.random.default_rng(12)
    D = rng.standard_normal((B, P, 3)); D = torch.from_numpy((D / np.linalg.norm(D, axis=-1, keepdims=True)).astype(np.float32)).to(dev)
    tg = torch.from_numpy(rng.uniform(-1, 1, (B, P, 3)).astype(np.float32)).to(dev)
    sw = torch.from_numpy(np.repeat(rng.uniform(0, 1, (B, P, 1)), 3, 2).astype(np.float32)).to(dev)
    Z = m.Z.detach()
    kw = dict(alpha=1e-3, beta=0.3 if cosine else 0.0, use_cosine=cosine, need_dw=True)
    a = F_.loss_forward_backward(m.spec, F_.Workspace(), Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), tile_major_bwd=True, **kw)
    torch.cuda.synchronize()
    for rep in range(6):
        b = F_.loss_forward_backward(m.spec, F_.Workspace(), Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), tile_major_bwd=False, **kw)
        torch.cuda.synchronize()
        print((B, P, N), "rep", rep, "dZ %.2e" % rel(b.dZ, a.dZ), "dW", " ".join("%.1e" % rel(x, y) for x, y in zip(b.dW, a.dW)),
              "db", " ".join("%.1e" % rel(x, y) for x, y in zip(b.db, a.db)))
