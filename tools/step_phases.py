"""Per-kernel times of one cfg-2 training step (library phase-event hook), both backward schedules.
    python tools/step_phases.py [maps] [reps]"""
import ctypes as C
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.build()
from reni_b200 import RENIAutoDecoder, _lib, get_directions, get_sineweight  # noqa: E402
from reni_b200 import functional as F_  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
W = 128
P = W * W // 2
m = RENIAutoDecoder(B, 36, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
tg = torch.rand(B, P, 3, device=dev) * 2 - 1
Z = m.Z.detach()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
n_ev = 7
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
for e in evs:
    e.record()
torch.cuda.synchronize()
handles = (C.c_void_p * n_ev)(*[e.cuda_event for e in evs])
names = {False: ["prologue", "fwd", "loss", "lbwd_head", "lbwd_layers", "map_level"],
         True: ["prologue", "fwd", "loss", "bwd_chain", "dw", "map_level"]}
for tile_major in (((os.environ.get("RENI_TILE_MAJOR_BWD") == "1"),) if os.environ.get("RENI_ONLY_LBWD") else (False, True)):
    ws = F_.Workspace()
    step = lambda: F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(),
                                            tile_major_bwd=tile_major)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.check(lib.reni_debug_set_phase_events(handles, n_ev))
    ms = [[] for _ in range(n_ev - 1)]
    for _ in range(reps):
        flush.zero_()
        step()
        torch.cuda.synchronize()
        for k in range(n_ev - 1):
            ms[k].append(evs[k].elapsed_time(evs[k + 1]))
    _lib.check(lib.reni_debug_set_phase_events(None, 0))
    tot = sum(statistics.median(v) for v in ms)
    print("tile-major" if tile_major else "layer-major", f"total {tot*1e3:.1f} us:",
          "  ".join(f"{n} {statistics.median(v)*1e3:.1f}" for n, v in zip(names[tile_major], ms)))
    # whole step, graph-free wall (events around the call)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("   eager step (forked streams):", f"{statistics.median(ts)*1e3:.1f} us")
