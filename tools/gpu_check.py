"""Developer diagnostics on a B200 (run under gpurun): tcgen05 descriptor self-tests, forward
parity against the golden fixtures, and a first timing.  Prints instead of asserting so one
GPU call yields as much information as possible; writes gpurun_out/gpu_check.json.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reni_oracle as O  # noqa: E402
from make_golden import CASES, golden_inputs  # noqa: E402

from reni_b200 import _lib  # noqa: E402

OUT = {}
dev = torch.device("cuda:0")
lib = _lib.load()


def vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def image_kmajor(mat):
    """[R x K] -> [K/8][R][8] fp16 image (numpy)."""
    R, K = mat.shape
    return np.ascontiguousarray(mat.astype(np.float16).reshape(R, K // 8, 8).transpose(1, 0, 2))


def selftest(name, A, Bm, a_img, b_img, a_lbo, a_sbo, b_lbo, b_sbo, a_k, b_k, a_mn, b_mn, N, ksteps, ref):
    ta = torch.from_numpy(a_img.view(np.uint8).reshape(-1)).to(dev)
    tb = torch.from_numpy(b_img.view(np.uint8).reshape(-1)).to(dev)
    d = torch.zeros(128, N, device=dev, dtype=torch.float32)
    rc = lib.reni_selftest_umma(vp(ta), ta.numel(), vp(tb), tb.numel(), a_lbo, a_sbo, b_lbo, b_sbo, a_k, b_k,
                                a_mn, b_mn, N, ksteps, vp(d), None)
    torch.cuda.synchronize()
    got = d.cpu().numpy()
    err = float(np.abs(got - ref).max())
    print(f"[selftest {name}] rc={rc} max|err|={err:.3e} ref_absmax={np.abs(ref).max():.3f}")
    OUT[f"selftest_{name}"] = err
    return got


def run_selftests():
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (128, 256)).astype(np.float16)
    Bw = rng.uniform(-1, 1, (256, 256)).astype(np.float16)
    ref = A.astype(np.float32) @ Bw.astype(np.float32).T
    # K-major A [k/8][128][8], B [k/8][256][8]; 16 k-steps, each 2 column groups
    selftest("kmajor_n256", A, Bw, image_kmajor(A), image_kmajor(Bw), 2048, 128, 4096, 128, 4096, 8192, 0, 0, 256, 16, ref)
    B16 = rng.uniform(-1, 1, (16, 256)).astype(np.float16)
    ref16 = A.astype(np.float32) @ B16.astype(np.float32).T
    selftest("kmajor_n16", A, B16, image_kmajor(A), image_kmajor(B16), 2048, 128, 256, 128, 4096, 512, 0, 0, 16, 16, ref16)
    # MN-major: operands are [64 rows x cols] half images; D[m,n] = sum_r At[r,m] Bt[r,n]
    At = rng.uniform(-1, 1, (64, 128)).astype(np.float16)
    Bt = rng.uniform(-1, 1, (64, 256)).astype(np.float16)
    refmn = At.astype(np.float32).T @ Bt.astype(np.float32)
    selftest("mnmajor", At, Bt, image_kmajor(At), image_kmajor(Bt), 128, 1024, 128, 1024, 256, 256, 1, 1, 256, 4, refmn)
    # same with LBO/SBO swapped, to learn the convention if the first guess is wrong
    selftest("mnmajor_swapped", At, Bt, image_kmajor(At), image_kmajor(Bt), 1024, 128, 1024, 128, 256, 256, 1, 1, 256, 4, refmn)


class Decoder:
    """Minimal driver of the C ABI (the product wrapper lives in reni_b200.models)."""

    def __init__(self, p: O.DecoderParams, N):
        self.p = p
        self.L = len(p.weights) - 2
        self.cfg = _lib.RENIConfig(N, _lib.EQUIVARIANCE[p.equivariance], 256, self.L, 3,
                                   1 if p.last_layer_linear else 0, 1 if p.output_activation == "tanh" else 0,
                                   p.first_omega_0, p.hidden_omega_0)
        self.w = [torch.from_numpy(w).to(dev).contiguous() for w in p.weights]
        self.b = [torch.from_numpy(b).to(dev).contiguous() for b in p.biases]

    def workspace(self, B, P, flags):
        n = lib.reni_workspace_bytes(C.byref(self.cfg), B, P, flags)
        assert n > 0, n
        ws = torch.empty(n + 1024, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 1024
        return ws[off:off + n], n

    def prepare(self, ws, n):
        nl = self.L + 2
        wp = (C.c_void_p * nl)(*[w.data_ptr() for w in self.w])
        bp = (C.c_void_p * nl)(*[b.data_ptr() for b in self.b])
        _lib.check(lib.reni_prepare_weights(C.byref(self.cfg), wp, bp, vp(ws), n, None), "prepare")

    def forward(self, Z, D, flags=0, target=None, sw=None, ws=None, n=None):
        B, P = Z.shape[0], D.shape[1]
        if ws is None:
            ws, n = self.workspace(B, P, flags)
            self.prepare(ws, n)
        out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
        dbs = 0 if D.shape[0] == 1 else P * 3
        sbs = 0 if (sw is None or sw.shape[0] == 1) else P * 3
        rc = lib.reni_forward(C.byref(self.cfg), vp(Z), vp(D), dbs, vp(self.w[0]), vp(self.b[0]), B, P, vp(out),
                              vp(target), vp(sw), sbs, vp(ws), n, flags, None)
        _lib.check(rc, "forward")
        return out, ws


def forward_parity():
    for name in ("so2_n9_h256", "so2_n36_h256", "so2_n36_h256_masked"):
        seed, B, P, N, H, L, out_f, eq, last_lin, act, grid, alpha, beta, full = CASES[name]
        p, Z, D, target, sw, mask = golden_inputs(seed, B, P, N, H, L, out_f, eq, grid_sidelen=grid)
        p.last_layer_linear, p.output_activation = last_lin, act
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        dec = Decoder(p, N)
        for flags in (0, 3):
            out, ws = dec.forward(torch.from_numpy(Z).to(dev), torch.from_numpy(D).to(dev), flags)
            torch.cuda.synchronize()
            o = out.cpu().numpy()
            e2, em = O.rel_l2(o, g["out_f32"]), O.rel_max(o, g["out_f32"])
            print(f"[forward {name} flags={flags}] rel_l2={e2:.3e} rel_max={em:.3e} finite={np.isfinite(o).all()}")
            OUT[f"fwd_{name}_{flags}"] = [e2, em]
        # prologue check: M_b, c_b against the oracle
        M, c = O.hoist_layer0(Z.astype(np.float64), p.weights[0].astype(np.float64), p.biases[0].astype(np.float64))
        cfg = dec.cfg
        # locate mc in workspace via a second tiny run is awkward; recompute offsets like abi.cu
        def al(x):
            return (x + 1023) // 1024 * 1024
        off = 0
        for sz in (L * 131072, L * 131072, 8192, 8192, (L * 256 + 16) * 4, 1024):
            off = al(off + sz)
        mc = ws[off:off + B * 5 * 256 * 4].view(torch.float32).cpu().numpy().reshape(B, 5, 256)
        ref_mc = np.concatenate((M, c[:, None, :]), 1) * p.first_omega_0
        print(f"[prologue {name}] rel_l2={O.rel_l2(mc, ref_mc):.3e}")
        OUT[f"prologue_{name}"] = O.rel_l2(mc, ref_mc)


def ptr_array(ts):
    return (C.c_void_p * len(ts))(*[(t.data_ptr() if t is not None else 0) for t in ts])


def fused_step(dec, Z, D, target, sw, alpha, beta, use_cos, need_dw, ws=None, n=None):
    B, P = Z.shape[0], D.shape[1]
    flags = 1 | 4 | (2 if need_dw else 0)
    if ws is None:
        ws, n = dec.workspace(B, P, flags)
        dec.prepare(ws, n)
    out = torch.empty(B, P, 3, device=dev)
    loss = torch.zeros(4, device=dev)
    dZ = torch.zeros_like(Z)
    dW = [torch.zeros_like(w) for w in dec.w] if need_dw else None
    db = [torch.zeros_like(b) for b in dec.b] if need_dw else None
    dbs = 0 if D.shape[0] == 1 else P * 3
    sbs = 0 if sw.shape[0] == 1 else P * 3
    rc = lib.reni_loss_forward_backward(
        C.byref(dec.cfg), vp(Z), vp(D), dbs, ptr_array(dec.w), ptr_array(dec.b), B, P, vp(target), vp(sw), sbs,
        alpha, beta, use_cos, vp(out), vp(loss), vp(dZ), ptr_array(dW) if need_dw else None,
        ptr_array(db) if need_dw else None, vp(ws), n, flags, None)
    _lib.check(rc, "loss_forward_backward")
    return out, loss, dZ, dW, db, ws, n


def backward_parity():
    from make_golden import sub_dw
    for name in ("so2_n9_h256", "so2_n36_h256", "so2_n36_h256_masked"):
        seed, B, P, N, H, L, out_f, eq, last_lin, act, grid, alpha, beta, full = CASES[name]
        p, Z, D, target, sw, mask = golden_inputs(seed, B, P, N, H, L, out_f, eq, grid_sidelen=grid)
        p.last_layer_linear, p.output_activation = last_lin, act
        if name.endswith("masked"):
            sw = sw * mask
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
        dec = Decoder(p, N)
        tZ, tD, tt, tsw = (torch.from_numpy(a).to(dev) for a in (Z, D, target, sw))
        # FIT_DECODER step: RENITrainLoss, all gradients
        out, loss, dZ, dW, db, ws, n = fused_step(dec, tZ, tD, tt, tsw, 0.0, 0.0, 0, True)
        torch.cuda.synchronize()
        print(f"[train {name}] loss={loss[0].item():.6f} ref={float(g['train_loss_f32']):.6f} "
              f"dZ rel_l2={O.rel_l2(dZ.cpu().numpy(), g['train_dZ_f32']):.3e}")
        OUT[f"train_{name}_dZ"] = O.rel_l2(dZ.cpu().numpy(), g["train_dZ_f32"])
        for i in range(L + 2):
            mine = dW[i].cpu().numpy()
            e = O.rel_l2(sub_dw(i, mine), g[f"train_dW{i}_f32"])
            nrm = float(np.linalg.norm(mine)) / float(g[f"train_dW{i}_norm_f32"])
            eb = O.rel_l2(db[i].cpu().numpy(), g[f"train_db{i}_f32"])
            print(f"   dW{i} rel_l2={e:.3e} norm_ratio={nrm:.5f}   db{i} rel_l2={eb:.3e}")
            OUT[f"train_{name}_dW{i}"] = [e, nrm, eb]
        # FIT_LATENT step: RENITestLoss, latent gradients only
        out, loss, dZ, _, _, ws, n = fused_step(dec, tZ, tD, tt, tsw, alpha, beta, 1, False)
        torch.cuda.synchronize()
        print(f"[latent {name}] loss={loss.cpu().numpy()} ref={g['test_loss_f32']} "
              f"dZ rel_l2={O.rel_l2(dZ.cpu().numpy(), g['test_dZ_f32']):.3e}")
        OUT[f"latent_{name}_dZ"] = O.rel_l2(dZ.cpu().numpy(), g["test_dZ_f32"])


def timing_step():
    rng = np.random.default_rng(1)
    N = 36
    p = O.siren_init(rng, N)
    dec = Decoder(p, N)
    for (B, W, need_dw) in ((32, 128, True), (32, 128, False), (256, 128, True)):
        P = W * W // 2
        Z = torch.randn(B, N, 3, device=dev)
        D = torch.from_numpy(O.get_directions(W)).to(dev)
        sw = torch.from_numpy(O.get_sineweight(W)).to(dev)
        target = torch.rand(B, P, 3, device=dev) * 2 - 1
        flags = 1 | 4 | (2 if need_dw else 0)
        ws, n = dec.workspace(B, P, flags)
        dec.prepare(ws, n)
        for _ in range(3):
            fused_step(dec, Z, D, target, sw, 0.0, 0.0, 0, need_dw, ws, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            fused_step(dec, Z, D, target, sw, 0.0, 0.0, 0, need_dw, ws, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rate = B * P / ms * 1e3
        fl = 1976832 if need_dw else 1317888
        print(f"[timing step B={B} P={P} need_dw={need_dw}] {ms:.3f} ms  {rate/1e6:.1f} M dirs/s  "
              f"{rate*fl/1e12:.1f} TFLOP/s ({rate*fl/1e12/1674.5*100:.1f}% of burst peak)")
        OUT[f"time_step_B{B}_dw{int(need_dw)}"] = [ms, rate]
        del ws


def timing():
    rng = np.random.default_rng(1)
    N = 36
    p = O.siren_init(rng, N)
    dec = Decoder(p, N)
    for (B, W) in ((32, 128), (256, 128)):
        P = W * W // 2
        Z = torch.randn(B, N, 3, device=dev)
        D = torch.from_numpy(O.get_directions(W)).to(dev)
        for flags in (0, 3):
            ws, n = dec.workspace(B, P, flags)
            dec.prepare(ws, n)
            for _ in range(3):
                dec.forward(Z, D, flags, ws=ws, n=n)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                dec.forward(Z, D, flags, ws=ws, n=n)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            rate = B * P / ms * 1e3
            tf = rate * 658944 / 1e12
            print(f"[timing fwd B={B} P={P} flags={flags}] {ms:.3f} ms  {rate/1e6:.1f} M dirs/s  {tf:.1f} TFLOP/s (fwd flops)")
            OUT[f"time_fwd_B{B}_f{flags}"] = [ms, rate]
            del ws


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "missing symbols:", _lib.missing_symbols())
    steps = sys.argv[1:] or ["selftest", "forward", "backward", "timing", "timing_step"]
    for s in steps:
        try:
            {"selftest": run_selftests, "forward": forward_parity, "timing": timing, "backward": backward_parity,
             "timing_step": timing_step}[s]()
        except Exception as e:  # keep going: one GPU call should tell us as much as possible
            import traceback

            traceback.print_exc()
            OUT[f"error_{s}"] = repr(e)
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                print("device unusable after error:", e2)
                break
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(OUT, f, indent=1)
