"""Time alternative builds of libreni_b200.so (build/variants/*.so, made by tools/build_variants.sh) on BASELINE
configs[1]: per-kernel CUDA-event times of the fused training step.  One subprocess per library (RENI_B200_LIB)."""
import ctypes as C, glob, os, statistics, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def child():
    import torch
    sys.path.insert(0, ROOT)
    from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight, _lib
    from reni_b200 import functional as F_
    mode = os.environ.get("RENI_MODE", "train")
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, N, W = int(os.environ.get("RENI_B", "32")), 36, 128
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, mode == "latent").to(dev)
    D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = torch.randn(B, N, 3, device=dev)
    ws = F_.Workspace()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def step():
        ws.prepared_key = None
        F_.loss_forward_backward(m.spec, ws, Z, D, tg, sw, m.decoder_weights(), m.decoder_biases(),
                                 alpha=0.0, beta=0.0, use_cosine=False, need_dw=mode == "train")
    lib = _lib.load()
    n_ev = 7
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
    for e in evs: e.record()
    torch.cuda.synchronize()
    handles = (C.c_void_p * n_ev)(*[e.cuda_event for e in evs])
    for _ in range(5): step()
    torch.cuda.synchronize()
    _lib.check(lib.reni_debug_set_phase_events(handles, n_ev))
    ph = [[] for _ in range(n_ev - 1)]
    for _ in range(int(os.environ.get("RENI_STEPS", "30"))):
        flush.zero_()
        step()
        torch.cuda.synchronize()
        for k in range(n_ev - 1): ph[k].append(evs[k].elapsed_time(evs[k + 1]))
    _lib.check(lib.reni_debug_set_phase_events(None, 0))
    names = ["prologue", "fwd", "loss", "bwd", "dw", "map"]
    med = [statistics.median(v) for v in ph]
    print(" ".join(f"{n}={v*1e3:.0f}us" for n, v in zip(names, med)), f"| total={sum(med)*1e3:.0f}us")

if __name__ == "__main__":
    if os.environ.get("RENI_CHILD"):
        child()
    else:
        libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
        for lib in libs:
            env = dict(os.environ, RENI_CHILD="1", RENI_B200_LIB=os.path.abspath(lib))
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
            out = (r.stdout.strip().splitlines() or ["(no output) " + r.stderr.strip()[-300:]])[-1]
            print(f"{os.path.basename(lib):40s} {out}", flush=True)
