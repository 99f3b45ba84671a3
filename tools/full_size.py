"""BASELINE configs 3, 4 and 5 at their FULL sizes on one B200 (the config sweep times shards / chunks of them):
  cfg3: N=100, 256 maps x 128x256, training step (all gradients)          -> 8.4 M directions, 52 GB of stash
  cfg4: frozen decoder, 4096 maps x 64x128, masked RENITestLoss, dZ only  -> 33.6 M directions, 103 GB of stash
  cfg5: inference, 1024 latents x 256x512, N=36                           -> 134 M directions, 1.6 GB of radiance
Each is timed (CUDA events, 3 runs after 1 warm-up) and checked for batch independence against a small-batch run of its
first and last maps (64-bit offsets through the whole stash).  Prints one JSON object."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.build()
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, rectangle_mask
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
OUT = {}

def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, r

def train(name, N, B, W, need_dw, **kw):
    torch.manual_seed(0)
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, not need_dw).to(dev)
    with torch.no_grad(): m.Z.normal_()
    D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
    if kw.get("mask") is not None: sw = sw * kw["mask"].to(dev)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = m.Z.detach()
    ws = F_.Workspace()
    args = dict(alpha=kw.get("alpha", 0.0), beta=kw.get("beta", 0.0), use_cosine=kw.get("use_cos", False), need_dw=need_dw)
    step = lambda: F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(), **args)
    ms, full = timeit(step)
    sel = [0, 1, B - 2, B - 1]
    small = F_.loss_forward_backward(m.spec, F_.Workspace(), Z[sel], D, tg[sel], sw, m.decoder_weights(), m.decoder_biases(), **args)
    torch.cuda.synchronize()
    ok_out = bool(torch.equal(full.out[sel], small.out))
    dz_err = float((full.dZ[sel] - small.dZ).norm() / small.dZ.norm())
    fl = 1976832 if need_dw else 1317888
    OUT[name] = {"maps": B, "directions": B * P, "ms_per_step": ms, "dirs_per_s": B * P / ms * 1e3,
                 "frac_bf16_peak": B * P * fl / (ms * 1e-3) / 1e12 / PEAK, "workspace_GiB": ws.nbytes / 2**30,
                 "first_last_maps_equal_small_batch": ok_out, "dZ_rel_diff_vs_small_batch": dz_err}
    print(name, OUT[name], flush=True)
    del ws, full, small, tg, m
    torch.cuda.empty_cache()

train("cfg3_full", 100, 256, 256, True)
train("cfg4_full", 36, 4096, 128, False, alpha=1e-7, beta=1e-4, use_cos=True, mask=rectangle_mask(128, 10, 46, 40, 82))

torch.manual_seed(0)
B, W, N = 1024, 512, 36
P = W * W // 2
m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D = get_directions(W).to(dev)
Z = m.Z.detach()
def f():
    with torch.no_grad(): return m(Z, D)
ms, o = timeit(f)
with torch.no_grad(): small = m(Z[[0, B - 1]], D)
OUT["cfg5_full_N36"] = {"latents": B, "directions": B * P, "ms": ms, "dirs_per_s": B * P / ms * 1e3,
                        "frac_bf16_peak": B * P * 658944 / (ms * 1e-3) / 1e12 / PEAK,
                        "first_last_maps_equal_small_batch": bool(torch.equal(o[[0, B - 1]], small))}
print(json.dumps(OUT, indent=1))
