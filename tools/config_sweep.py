"""Timing (and a spot parity check) of the other BASELINE configs on one B200: cfg3 shape (N=100, 128x256), cfg4
(latent-only, masked RENITestLoss), cfg5 (inference sweep over N at 256x512).  Run under gpurun."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import reni_oracle as O
from helpers import params_from_model
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, rectangle_mask
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
OUT = {}

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def train_cfg(name, N, B, W, need_dw, alpha=0.0, beta=0.0, use_cos=False, mask=None, check=True):
    torch.manual_seed(0)
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, not need_dw).to(dev)
    if not need_dw:
        with torch.no_grad(): m.Z.normal_()
    D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
    if mask is not None: sw = sw * mask.to(dev)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = m.Z.detach()
    ws = F_.Workspace()
    def step():
        ws.prepared_key = None
        return F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), alpha=alpha, beta=beta, use_cosine=use_cos, need_dw=need_dw)
    ms = timeit(step)
    rate = B * P / ms * 1e3
    fl = 1976832 if need_dw else 1317888
    msg = f"[{name}] N={N} B={B} {W//2}x{W} need_dw={need_dw}: {ms:.3f} ms/step, {rate/1e6:.1f} M dirs/s, {rate*fl/1e12:.0f} TFLOP/s ({rate*fl/1e12/1674.5*100:.1f}% of burst peak), ws {ws.nbytes/2**30:.2f} GiB"
    if check:
        r = step(); torch.cuda.synchronize()
        b = B // 2
        p64 = params_from_model(m)
        # one map against the oracle, on a pixel subset to bound the N^2 encoding cost
        sel = slice(0, P, max(1, P // 2048))
        Zb = Z[[b]].cpu().numpy().astype(np.float64)
        o_ref = O.decoder_forward(Zb, D.cpu().numpy().astype(np.float64)[:, sel], p64)
        msg += f" | radiance rel-L2 vs oracle (1 map, subset) {O.rel_l2(r.out[[b]][:, sel].cpu().numpy(), o_ref):.2e}"
    print(msg, flush=True); OUT[name] = [ms, rate]

def infer_cfg(N, B=16, W=512):
    torch.manual_seed(0)
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    D = get_directions(W).to(dev)
    Z = m.Z.detach()
    def f():
        with torch.no_grad(): return m(Z, D)
    ms = timeit(f)
    rate = B * P / ms * 1e3
    print(f"[cfg5 infer] N={N} B={B} {W//2}x{W}: {ms:.3f} ms, {rate/1e6:.1f} M dirs/s, {rate*658944/1e12:.0f} TFLOP/s ({rate*658944/1e12/1674.5*100:.1f}% of burst peak)", flush=True)
    OUT[f"cfg5_N{N}"] = [ms, rate]

train_cfg("cfg2", 36, 32, 128, True)
train_cfg("cfg3-shard (32 of 256 maps)", 100, 32, 256, True)
train_cfg("cfg4-chunk (256 of 4096 maps)", 36, 256, 128, False, alpha=1e-7, beta=1e-4, use_cos=True, mask=rectangle_mask(128, 10, 46, 40, 82))
for N in (9, 36, 49, 100): infer_cfg(N)
json.dump(OUT, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
