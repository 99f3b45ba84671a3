"""Generate tests/golden/film_*.npz by running the UNMODIFIED reference FiLM decoder
(JADGardner/RENI ``RENIAutoDecoderFiLM``, src/models/RENI.py:527-678).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_film.py        (build container: reference mounted at /root/reference)

Inputs come from ``film_golden_inputs`` (numpy RNG, shared with the tests); the fixtures hold the reference's
fp32 outputs and an fp64 "truth" from the same modules after ``.double()``.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)

from make_golden import _import_reference, unit  # noqa: E402

# name: (seed, B, P, N, H, Lf, map_features, map_layers, out_f, equiv, out_act, grid_sidelen, alpha, beta, full)
FILM_CASES = {
    "film_so2_small": (21, 3, 48, 4, 32, 3, 16, 2, 3, "SO2", None, 0, 1e-3, 0.5, True),
    "film_so3_small_tanh": (22, 2, 40, 5, 32, 2, 24, 1, 3, "SO3", "tanh", 0, 1e-3, 0.5, True),
    "film_so2_small_exp": (23, 2, 32, 3, 16, 2, 16, 2, 3, "SO2", "exp", 0, 1e-7, 1e-4, True),
    "film_so2_n9_h256": (24, 2, 128, 9, 256, 5, 256, 3, 3, "SO2", None, 16, 1e-7, 1e-4, False),
    "film_so2_n36_h256": (25, 3, 512, 36, 256, 5, 256, 3, 3, "SO2", None, 32, 1e-7, 1e-4, False),
    "film_so3_n9_h256_tanh": (26, 2, 128, 9, 256, 3, 256, 3, 3, "SO3", "tanh", 16, 1e-7, 1e-1, False),
}


def film_golden_inputs(seed, B, P, N, H, Lf, map_f, map_l, out_f, eq, act, grid_sidelen=0):
    import reni_film_oracle as FO
    import reni_oracle as O

    rng = np.random.default_rng(seed)
    p = FO.film_init(rng, N, eq, H, Lf, map_f, map_l, out_f, act)
    # a trained mapping network moves freq / phase away from their init (freq ~ 30, phase ~ 0): scale its last
    # layer up so the per-map modulation is exercised
    p.map_w[-1] = (2.0 * p.map_w[-1]).astype(np.float32)
    Z = (0.5 * rng.standard_normal((B, N, 3))).astype(np.float32)
    if grid_sidelen:
        D = np.repeat(O.get_directions(grid_sidelen), B, 0)
        sw = np.repeat(O.get_sineweight(grid_sidelen), B, 0)
        assert D.shape[1] == P
    else:
        D = unit(rng.standard_normal((B, P, 3))).astype(np.float32)
        sw = np.repeat(rng.uniform(0.0, 1.0, (B, P, 1)), 3, 2).astype(np.float32)
    target = rng.uniform(-1, 1, (B, P, out_f)).astype(np.float32)
    return p, Z, D, target, sw


def _make_ref_model(ref_model, p, N, dtype):
    import torch

    Lf, H = len(p.net_w), p.net_w[0].shape[0]
    m = ref_model.RENIAutoDecoderFiLM(4, N, p.equivariance, H, Lf, p.map_w[0].shape[0], len(p.map_w) - 1,
                                      p.final_w.shape[0], p.output_activation, False)
    sd = {}
    for i in range(Lf):
        sd[f"net.{i}.layer.weight"] = torch.from_numpy(p.net_w[i])
        sd[f"net.{i}.layer.bias"] = torch.from_numpy(p.net_b[i])
    sd["final_layer.weight"] = torch.from_numpy(p.final_w)
    sd["final_layer.bias"] = torch.from_numpy(p.final_b)
    for i in range(len(p.map_w)):
        sd[f"mapping_network.network.{2 * i}.weight"] = torch.from_numpy(p.map_w[i])
        sd[f"mapping_network.network.{2 * i}.bias"] = torch.from_numpy(p.map_b[i])
    sd["Z"] = m.Z.detach().clone()
    torch.nn.Module.load_state_dict(m, sd, strict=True)
    return m.to(dtype)


def run_case(ref_model, ref_loss, p, Z, D, target, sw, dtype, alpha, beta):
    import torch

    tdt = torch.float32 if dtype == np.float32 else torch.float64
    model = _make_ref_model(ref_model, p, Z.shape[1], tdt)
    tZ = torch.tensor(Z, dtype=tdt, requires_grad=True)
    tD, tt, tsw = (torch.tensor(a, dtype=tdt) for a in (D, target, sw))
    out = model(tZ, tD)
    loss = ref_loss.RENITrainLoss()(out, tt, tsw)
    loss.backward()
    r = {"out": out.detach().numpy(), "train_loss": loss.detach().numpy(), "train_dZ": tZ.grad.numpy().copy()}
    Lf = len(p.net_w)
    for i in range(Lf):
        r[f"net_dW{i}"] = model.net[i].layer.weight.grad.numpy().copy()
        r[f"net_db{i}"] = model.net[i].layer.bias.grad.numpy().copy()
    r["final_dW"] = model.final_layer.weight.grad.numpy().copy()
    r["final_db"] = model.final_layer.bias.grad.numpy().copy()
    for i in range(len(p.map_w)):
        r[f"map_dW{i}"] = model.mapping_network.network[2 * i].weight.grad.numpy().copy()
        r[f"map_db{i}"] = model.mapping_network.network[2 * i].bias.grad.numpy().copy()
    # latent fit (RENITestLoss) on the same model
    model.zero_grad()
    tZ2 = torch.tensor(Z, dtype=tdt, requires_grad=True)
    out2 = model(tZ2, tD)
    l, mse, prior, cos = ref_loss.RENITestLoss(alpha=alpha, beta=beta)(out2, tt, tsw, tZ2)
    l.backward()
    r["test_loss"] = np.array([l.item(), mse.item(), prior.item(), cos.item()])
    r["test_dZ"] = tZ2.grad.numpy().copy()
    return r


def sub(w):
    """Strided subsample of the big gradient matrices (full tensors would be MBs per case)."""
    if w.ndim == 2 and w.shape[0] * w.shape[1] > 40000:
        return w[:: max(1, w.shape[0] // 32), :: max(1, w.shape[1] // 32)]
    return w


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref_model, ref_loss, _ = _import_reference()
    import torch

    torch.set_num_threads(8)
    for name, (seed, B, P, N, H, Lf, mf, ml, out_f, eq, act, grid, alpha, beta, full) in FILM_CASES.items():
        p, Z, D, target, sw = film_golden_inputs(seed, B, P, N, H, Lf, mf, ml, out_f, eq, act, grid)
        store = {}
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            r = run_case(ref_model, ref_loss, p, Z, D, target, sw, dt, alpha, beta)
            for k, v in r.items():
                if k.startswith(("net_dW", "map_dW")):
                    store[f"{k}_norm_{tag}"] = np.array(np.linalg.norm(v.astype(np.float64)))
                    v = v if full else sub(v)
                store[f"{k}_{tag}"] = v
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **store)
        print(name, store["out_f32"].shape, float(np.abs(store["out_f32"]).mean()))

    # module-level behaviour of the FiLM classes: state-dict keys / shapes / init ranges, dispatch
    torch.manual_seed(0)
    m = ref_model.RENIAutoDecoderFiLM(7, 36, "SO2", 256, 5, 256, 3, 3, None, False)
    mod = {}
    names, shapes, los, his = [], [], [], []
    for k, v in m.state_dict().items():
        names.append(k)
        shapes.append(list(v.shape) + [0] * (3 - v.dim()))
        los.append(float(v.min()))
        his.append(float(v.max()))
    mod["state_keys"] = np.array(names)
    mod["state_shapes"] = np.array(shapes, dtype=np.int64)
    mod["state_min"] = np.array(los)
    mod["state_max"] = np.array(his)
    mod["state_std"] = np.array([float(v.float().std()) if v.numel() > 1 else 0.0 for v in m.state_dict().values()])
    mv = ref_model.RENIVariationalAutoDecoderFiLM(7, 36, "SO2", 256, 5, 256, 3, 3, None, True)
    mod["vad_state_keys"] = np.array(list(mv.state_dict().keys()))
    mod["vad_fixed_requires_grad"] = np.array([int(q.requires_grad) for q in mv.parameters()])
    mod["vad_fixed_mu_absmax"] = np.array(float(mv.mu.abs().max()))
    # dispatch (RENI.py:626-664) on a small model
    torch.manual_seed(1)
    ms = ref_model.RENIAutoDecoderFiLM(6, 4, "SO2", 32, 2, 16, 1, 3, "tanh", False)
    Dd = torch.from_numpy(unit(np.random.default_rng(5).standard_normal((2, 16, 3))).astype(np.float32))
    for i, (k, v) in enumerate(ms.state_dict().items()):
        mod[f"disp_param_{i}"] = v.numpy()
    mod["disp_param_keys"] = np.array(list(ms.state_dict().keys()))
    mod["disp_D"] = Dd.numpy()
    mod["disp_out_int"] = ms(3, Dd[:1]).detach().numpy()
    mod["disp_out_list"] = ms([1, 4], Dd).detach().numpy()
    mod["disp_out_idx"] = ms(torch.tensor([5, 0]), Dd).detach().numpy()
    mod["disp_out_lat"] = ms(ms.Z[[2, 3]], Dd).detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "film_module.npz"), **mod)
    print("FiLM golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
