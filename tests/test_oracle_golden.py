"""Pin the oracle (oracle/reni_oracle.py) against fixtures produced by the reference itself
(tests/golden/*.npz, written by oracle/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest

import reni_oracle as O
from make_golden import CASES, golden_inputs, sub_dw

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_case(name):
    seed, B, P, N, H, L, out_f, eq, last_lin, act, grid, alpha, beta, full = CASES[name]
    p, Z, D, target, sw, mask = golden_inputs(seed, B, P, N, H, L, out_f, eq, grid_sidelen=grid)
    p.last_layer_linear = last_lin
    p.output_activation = act
    if name.endswith("masked"):
        sw = sw * mask
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return p, Z, D, target, sw, g, alpha, beta, full


@pytest.mark.parametrize("name", list(CASES))
def test_forward_and_losses_match_reference(name):
    p, Z, D, target, sw, g, alpha, beta, full = load_case(name)
    # fp64 oracle vs fp64 reference: only reduction-order noise
    p64 = p.astype(np.float64)
    o64 = O.decoder_forward(Z.astype(np.float64), D.astype(np.float64), p64)
    assert O.rel_max(o64, g["out_f64"]) < 1e-11
    # the fp32 reference sits ~1e-6 from truth
    assert O.rel_max(g["out_f32"], o64) < 2e-5
    # fp32 oracle vs fp32 reference
    o32 = O.decoder_forward(Z, D, p)
    assert o32.dtype == np.float32
    assert O.rel_max(o32, g["out_f32"]) < 2e-5
    t64, s64 = target.astype(np.float64), sw.astype(np.float64)
    assert abs(O.reni_train_loss(o64, t64, s64) - g["train_loss_f64"]) < 1e-12 * max(1, abs(g["train_loss_f64"]))
    tl = O.reni_test_loss(o64, t64, s64, Z.astype(np.float64), alpha, beta)
    np.testing.assert_allclose(np.array(tl), g["test_loss_f64"], rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("name", list(CASES))
def test_backward_matches_reference_autograd(name):
    p, Z, D, target, sw, g, alpha, beta, full = load_case(name)
    p64 = p.astype(np.float64)
    Z64, D64, t64, s64 = (a.astype(np.float64) for a in (Z, D, target, sw))
    r = O.step_fit_decoder(Z64, D64, t64, s64, p64)
    assert O.rel_l2(r["dZ"], g["train_dZ_f64"]) < 1e-10
    for i, (dw, db) in enumerate(zip(r["dW"], r["db"])):
        ref = g[f"train_dW{i}_f64"]
        mine = dw if full else sub_dw(i, dw)
        assert O.rel_l2(mine, ref) < 1e-10, f"dW{i}"
        assert abs(np.linalg.norm(dw) - g[f"train_dW{i}_norm_f64"]) < 1e-9 * g[f"train_dW{i}_norm_f64"]
        assert O.rel_l2(db, g[f"train_db{i}_f64"]) < 1e-10, f"db{i}"
    r2 = O.step_fit_latent(Z64, D64, t64, s64, p64, alpha, beta)
    assert O.rel_l2(r2["dZ"], g["test_dZ_f64"]) < 1e-10
    # fp32 reference gradients agree with fp64 truth far inside the 1e-2 tolerance
    assert O.rel_l2(g["train_dZ_f32"], g["train_dZ_f64"]) < 1e-4


@pytest.mark.parametrize("name", list(CASES))
def test_encoding_matches_reference(name):
    p, Z, D, target, sw, g, alpha, beta, full = load_case(name)
    e = O.ENCODINGS[p.equivariance](Z, D)
    assert e.shape[-1] == O.in_features(Z.shape[1], p.equivariance) == p.weights[0].shape[1]
    cs = np.array([e.astype(np.float64).sum(), np.abs(e.astype(np.float64)).sum()])
    np.testing.assert_allclose(cs, g["enc_checksum"], rtol=1e-6)
    if full:
        np.testing.assert_allclose(e, g["enc"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("name", [n for n in CASES if CASES[n][7] == "SO2"] + ["so3_small", "none_small"])
def test_hoisted_equals_direct(name):
    p, Z, D, target, sw, g, alpha, beta, full = load_case(name)
    p64 = p.astype(np.float64)
    Z64, D64 = Z.astype(np.float64), D.astype(np.float64)
    o_direct = O.decoder_forward(Z64, D64, p64)
    o_hoist = O.hoisted_forward(Z64, D64, p64)
    assert O.rel_max(o_hoist, o_direct) < 1e-11


@pytest.mark.parametrize("name", ["so2_small", "so2_n9_h256", "so2_n36_h256_masked"])
def test_layer0_map_backward(name):
    """dM_b/dc_b -> (dW0, db0, dZ) equals the direct backward."""
    p, Z, D, target, sw, g, alpha, beta, full = load_case(name)
    p64 = p.astype(np.float64)
    Z64, D64, t64, s64 = (a.astype(np.float64) for a in (Z, D, target, sw))
    o, tape = O.decoder_forward(Z64, D64, p64, tape=True)
    go = O.loss_grad_wrt_output(o, t64, s64)
    dWs, dbs, dZ = O.decoder_backward(Z64, D64, p64, tape, go)
    # recover delta0 (gradient wrt omega-free pre-activation of layer 0)
    gg = go * (1 - o**2) if p64.output_activation == "tanh" else go
    nl = len(p64.weights)
    for i in reversed(range(1, nl)):
        is_sine = (i < nl - 1) or (not p64.last_layer_linear)
        if is_sine:
            gg = gg * np.cos(tape.pre[i]) * p64.hidden_omega_0
        gg = gg @ p64.weights[i]
    delta0 = gg * np.cos(tape.pre[0]) * p64.first_omega_0
    f = O.direction_features(D64)
    dc = delta0.sum(1)
    dM = np.einsum("bpf,bph->bfh", f, delta0)
    dW0, db0, dZ2 = O.layer0_backward_so2(Z64, p64.weights[0], dM, dc)
    assert O.rel_l2(dW0, dWs[0]) < 1e-11
    assert O.rel_l2(db0, dbs[0]) < 1e-11
    assert O.rel_l2(dZ2, dZ) < 1e-11


def test_so2_invariance_property():
    """model(Z R^T, D R^T) == model(Z, D) for rotations about y."""
    p, Z, D, *_ = load_case("so2_small")
    th = 0.7
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    p64 = p.astype(np.float64)
    a = O.decoder_forward(Z.astype(np.float64), D.astype(np.float64), p64)
    b = O.decoder_forward(Z.astype(np.float64) @ R.T, D.astype(np.float64) @ R.T, p64)
    assert np.abs(a - b).max() < 1e-12


def test_geometry_matches_reference():
    g = np.load(os.path.join(GOLDEN, "geometry.npz"))
    for W in (16, 32):
        np.testing.assert_allclose(O.get_directions(W), g[f"dir_{W}"], atol=2e-7)
        np.testing.assert_allclose(O.get_sineweight(W), g[f"sw_{W}"], atol=2e-7)
    d = O.get_directions(128).astype(np.float64)
    np.testing.assert_allclose([d.sum(), np.abs(d).sum(), (d**2).sum()], g["dir_128_checksum"], rtol=1e-6, atol=1e-3)
    s = O.get_sineweight(128).astype(np.float64)
    np.testing.assert_allclose([s.sum(), (s**2).sum()], g["sw_128_checksum"], rtol=1e-6)
    # |d_xz| == sineweight on the grid, |d| == 1
    d32 = O.get_directions(32).astype(np.float64)
    np.testing.assert_allclose(np.sqrt(d32[..., 0] ** 2 + d32[..., 2] ** 2), O.get_sineweight(32)[..., 0], atol=3e-7)
    np.testing.assert_allclose(np.linalg.norm(d32, axis=-1), 1.0, atol=3e-7)


def test_kld_matches_reference():
    g = np.load(os.path.join(GOLDEN, "module.npz"))
    v = O.kld(g["kld_mu"].astype(np.float64), g["kld_lv"].astype(np.float64), Z_dims=12)
    assert abs(v - g["kld_val"]) < 1e-5 * abs(g["kld_val"])


def test_init_distribution_matches_reference_constructor():
    """siren_init draws from the same ranges as RENI.py:76-84,157-160 (not the same stream)."""
    g = np.load(os.path.join(GOLDEN, "module.npz"))
    p = O.siren_init(np.random.default_rng(0), 36)
    keys = list(g["state_keys"])
    for i in range(7):
        kw = "net.%d.linear.weight" % i if i < 6 else "net.6.weight"
        j = keys.index(kw)
        assert list(g["state_shapes"][j][:2]) == list(p.weights[i].shape)
        lim = max(abs(g["state_min"][j]), abs(g["state_max"][j]))
        mine = np.abs(p.weights[i]).max()
        assert 0.9 * lim < mine < 1.1 * lim
    assert int(g["n_net_params"]) == sum(w.size for w in p.weights) + sum(b.size for b in p.biases) == 680707


def test_torch_port_matches_reference_golden():
    """oracle/reni_torch_port.py (the CPU baseline that bench.py times) == the reference's fp32 results."""
    import torch

    import reni_torch_port as TP

    p, Z, D, target, sw, g, alpha, beta, full = load_case("so2_n9_h256")
    ws = [torch.from_numpy(w) for w in p.weights]
    bs = [torch.from_numpy(b) for b in p.biases]
    loss, out, grads = TP.training_step(torch.from_numpy(Z), torch.from_numpy(D), torch.from_numpy(target),
                                        torch.from_numpy(sw), ws, bs)
    assert O.rel_max(out.numpy(), g["out_f32"]) < 2e-5
    assert abs(float(loss) - float(g["train_loss_f32"])) < 1e-5 * abs(float(g["train_loss_f32"]))
    assert O.rel_l2(grads[0].numpy(), g["train_dZ_f32"]) < 1e-4
    nl = len(ws)
    for i in range(nl):
        assert O.rel_l2(sub_dw(i, grads[1 + i].numpy()), g[f"train_dW{i}_f32"]) < 1e-4
        assert O.rel_l2(grads[1 + nl + i].numpy(), g[f"train_db{i}_f32"]) < 1e-4
