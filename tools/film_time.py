"""FiLM decoder timings on a B200 (N=36, default FiLM config: 5 FiLM layers, 3x256 mapping network):
  * inference: latents x directions decode sweep (CUDA events, L2 flushed between runs)
  * training step through the public API (RENITrainer on RENIAutoDecoderFiLM: autograd over the fused core), with the
    per-kernel split from the library's phase-event hook.
Prints one JSON object; `python tools/film_time.py > profiles/<name>.json`."""
import ctypes as C, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry

def main():
    entry.build()
    from reni_b200 import RENIAutoDecoderFiLM, RENITrainer, _lib, get_directions
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    torch.manual_seed(0)
    N, Lf = 36, 5
    L = Lf - 1
    flops_fwd = 2 * (4 * 256 + L * 256 * 256 + 256 * 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {"config": {"N": N, "film_layers": Lf, "mapping": "3 x 256"}, "flops_per_direction_fwd": flops_fwd}

    # ---- inference sweep
    inf = {}
    for B, W in ((32, 128), (256, 256), (1024, 256)):
        P = W * W // 2
        m = RENIAutoDecoderFiLM(B, N, "SO2", 256, Lf, 256, 3, 3, None, True).to(dev)
        D = get_directions(W).to(dev).expand(B, -1, -1)
        Z = torch.randn(B, N, 3, device=dev)
        with torch.no_grad():
            for _ in range(3): m(Z, D)
            ts = []
            for _ in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); m(Z, D); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        inf[f"{B}x{W//2}x{W}"] = {"ms": round(ms, 4), "dirs_per_s": B * P / ms * 1e3,
                                  "frac_bf16_peak": flops_fwd * B * P / (ms * 1e-3) / 1e12 / peak}
        del m
    res["inference"] = inf

    # ---- training step, cfg-2 shape: 32 maps x 64x128
    B, W = 32, 128
    P = W * W // 2
    for task in ("FIT_DECODER", "FIT_LATENT"):
        m = RENIAutoDecoderFiLM(B, N, "SO2", 256, Lf, 256, 3, 3, None, task == "FIT_LATENT").to(dev)
        if task == "FIT_LATENT":
            with torch.no_grad(): m.Z.normal_()
        tr = RENITrainer(m, task, W, lr=1e-5)
        imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
        idx = torch.arange(B, device=dev)
        for _ in range(5): tr.training_step((imgs, idx))
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.training_step((imgs, idx)); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        lib = _lib.load()
        n_ev = 7
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
        for e in evs: e.record()
        torch.cuda.synchronize()
        handles = (C.c_void_p * n_ev)(*[e.cuda_event for e in evs])
        ph = {"fwd_kernel": [], "bwd_kernel": [], "dw+film_reduce": []}
        _lib.check(lib.reni_debug_set_phase_events(handles, n_ev))
        for _ in range(20):
            flush.zero_()
            tr.training_step((imgs, idx)); torch.cuda.synchronize()
            ph["fwd_kernel"].append(evs[1].elapsed_time(evs[2]))
            ph["bwd_kernel"].append(evs[3].elapsed_time(evs[4]))
            ph["dw+film_reduce"].append(evs[4].elapsed_time(evs[5]))
        _lib.check(lib.reni_debug_set_phase_events(None, 0))
        trg = RENITrainer(m, task, W, lr=1e-5, cuda_graph=True)
        for _ in range(3): trg.training_step((imgs, idx))
        torch.cuda.synchronize()
        tg_ = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); trg.training_step((imgs, idx)); e1.record(); torch.cuda.synchronize()
            tg_.append(e0.elapsed_time(e1))
        ms_eager, ms = ms, statistics.median(tg_)
        mult = 3 if task == "FIT_DECODER" else 2
        res[task] = {"ms_per_step": round(ms, 4), "ms_per_step_eager": round(ms_eager, 4), "dirs_per_s": B * P / ms * 1e3,
                     "frac_bf16_peak_algorithmic": mult * flops_fwd * B * P / (ms * 1e-3) / 1e12 / peak,
                     "kernels_ms": {k: round(statistics.median(v), 4) for k, v in ph.items()}}
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
