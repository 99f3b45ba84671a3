"""FiLM-conditioned decoder on the GPU (run with -m gpu on a B200): the CUDA core behind RENIAutoDecoderFiLM, called
through the drop-in module (autograd -> reni_film_forward / reni_film_backward), against (a) golden fixtures produced by
the reference's RENIAutoDecoderFiLM and (b) the numpy FiLM oracle on seeded inputs.  Tolerances as for the
Cond-by-Concat path: radiance 1e-3 relative, gradients 1e-2 relative (rel-L2)."""
import numpy as np
import pytest
import torch

from helpers import O, TOL_GRAD, TOL_RADIANCE, TOL_RADIANCE_MAX

import reni_film_oracle as FO  # noqa: E402
from make_golden_film import sub  # noqa: E402
from test_film_models_cpu import film_model_from_params
from test_film_oracle_golden import load_film_case

pytestmark = pytest.mark.gpu

FILM_H256 = ["film_so2_n9_h256", "film_so2_n36_h256", "film_so3_n9_h256_tanh"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import __graft_entry__ as entry

    entry.build()
    return torch.device("cuda:0")


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("name", FILM_H256)
def test_film_forward_matches_reference_golden(dev, name):
    c = load_film_case(name)
    m = film_model_from_params(c["p"], c["N"], dev)
    with torch.no_grad():
        out = m(t(c["Z"], dev), t(c["D"], dev)).cpu().numpy()
    ref = c["g"]["out_f32"]
    assert out.shape == ref.shape
    assert O.rel_l2(out, ref) < TOL_RADIANCE
    assert O.rel_max(out, ref) < TOL_RADIANCE_MAX


@pytest.mark.parametrize("name", FILM_H256)
def test_film_training_and_latent_gradients_match_reference_golden(dev, name):
    """The reference's calling pattern: out = model(Z, D); loss = criterion(out, ...); loss.backward()."""
    from reni_b200 import RENITestLoss, RENITrainLoss

    c = load_film_case(name)
    g = c["g"]
    m = film_model_from_params(c["p"], c["N"], dev)
    Z = t(c["Z"], dev).requires_grad_(True)
    D, tg, sw = (t(c[k], dev) for k in ("D", "target", "sw"))
    out = m(Z, D)
    loss = RENITrainLoss()(out, tg, sw)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g["train_loss_f32"])) < 1e-3 * abs(float(g["train_loss_f32"]))
    assert O.rel_l2(out.detach().cpu().numpy(), g["out_f32"]) < TOL_RADIANCE
    assert O.rel_l2(Z.grad.cpu().numpy(), g["train_dZ_f32"]) < TOL_GRAD
    for i in range(c["Lf"]):
        dw = m.net[i].layer.weight.grad.cpu().numpy()
        assert O.rel_l2(sub(dw), g[f"net_dW{i}_f32"]) < TOL_GRAD, f"net_dW{i}"
        assert abs(np.linalg.norm(dw) / float(g[f"net_dW{i}_norm_f32"]) - 1) < TOL_GRAD
        assert O.rel_l2(m.net[i].layer.bias.grad.cpu().numpy(), g[f"net_db{i}_f32"]) < TOL_GRAD, f"net_db{i}"
    assert O.rel_l2(m.final_layer.weight.grad.cpu().numpy(), g["final_dW_f32"]) < TOL_GRAD
    assert O.rel_l2(m.final_layer.bias.grad.cpu().numpy(), g["final_db_f32"]) < TOL_GRAD
    for i in range(len(c["p"].map_w)):
        lin = m.mapping_network.network[2 * i]
        assert O.rel_l2(sub(lin.weight.grad.cpu().numpy()), g[f"map_dW{i}_f32"]) < TOL_GRAD, f"map_dW{i}"
        assert O.rel_l2(lin.bias.grad.cpu().numpy(), g[f"map_db{i}_f32"]) < TOL_GRAD, f"map_db{i}"

    # FIT_LATENT: frozen decoder, RENITestLoss (prior + cosine), latent gradients only
    mf = film_model_from_params(c["p"], c["N"], dev, fixed=True)
    Z2 = t(c["Z"], dev).requires_grad_(True)
    out2 = mf(Z2, D)
    l, mse, prior, cos = RENITestLoss(alpha=c["alpha"], beta=c["beta"])(out2, tg, sw, Z2)
    l.backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose([float(l), float(mse), float(prior), float(cos)], g["test_loss_f32"], rtol=1e-3, atol=1e-8)
    assert O.rel_l2(Z2.grad.cpu().numpy(), g["test_dZ_f32"]) < TOL_GRAD
    assert all(p.grad is None for p in mf.core_parameters())


@pytest.mark.parametrize("eq,act,P,B,Lf", [("SO2", "exp", 200, 5, 5), ("SO3", None, 1024, 3, 2), ("SO2", "tanh", 8192, 2, 7)])
def test_film_vs_oracle_ragged_tiles_exp_and_depth(dev, eq, act, P, B, Lf):
    """Shapes the fixtures do not cover: a ragged last tile (P = 200), several maps per CTA pair, the exp output
    activation, 2 and 7 FiLM layers, 64 tiles per map."""
    rng = np.random.default_rng(101)
    N = 7
    p = FO.film_init(rng, N, eq, 256, Lf, 64, 2, 3, act)
    p.map_w[-1] = (2.0 * p.map_w[-1]).astype(np.float32)
    Z = (0.5 * rng.standard_normal((B, N, 3))).astype(np.float32)
    Dn = rng.standard_normal((B, P, 3))
    D = (Dn / np.linalg.norm(Dn, axis=-1, keepdims=True)).astype(np.float32)
    tg = rng.uniform(-1, 1, (B, P, 3)).astype(np.float32)
    sw = np.repeat(rng.uniform(0, 1, (B, P, 1)), 3, 2).astype(np.float32)
    p64 = p.astype(np.float64)
    out_o, tape = FO.film_forward(Z.astype(np.float64), D.astype(np.float64), p64, tape=True)
    go = O.loss_grad_wrt_output(out_o, tg.astype(np.float64), sw.astype(np.float64))
    ref = FO.film_backward(Z.astype(np.float64), D.astype(np.float64), p64, tape, go)

    from reni_b200 import RENITrainLoss

    m = film_model_from_params(p, N, dev)
    Zt = t(Z, dev).requires_grad_(True)
    out = m(Zt, t(D, dev))
    RENITrainLoss()(out, t(tg, dev), t(sw, dev)).backward()
    torch.cuda.synchronize()
    assert O.rel_l2(out.detach().cpu().numpy(), out_o) < TOL_RADIANCE
    assert O.rel_l2(Zt.grad.cpu().numpy(), ref["dZ"]) < TOL_GRAD
    for i in range(Lf):
        assert O.rel_l2(m.net[i].layer.weight.grad.cpu().numpy(), ref["net_dW"][i]) < TOL_GRAD, i
        assert O.rel_l2(m.net[i].layer.bias.grad.cpu().numpy(), ref["net_db"][i]) < TOL_GRAD, i
    assert O.rel_l2(m.final_layer.weight.grad.cpu().numpy(), ref["final_dW"]) < TOL_GRAD
    for i in range(len(p.map_w)):
        assert O.rel_l2(m.mapping_network.network[2 * i].weight.grad.cpu().numpy(), ref["map_dW"][i]) < TOL_GRAD, i


def test_film_core_gradients_wrt_film_and_mc(dev):
    """The C-ABI core in isolation: d_mc and d_film against the oracle's dfreq / dphase."""
    from reni_b200 import functional as F_

    c = load_film_case("film_so2_n9_h256")
    p64 = c["p"].astype(np.float64)
    Z, D, tg, sw = (c[k].astype(np.float64) for k in ("Z", "D", "target", "sw"))
    out_o, tape = FO.film_forward(Z, D, p64, tape=True)
    go = O.loss_grad_wrt_output(out_o, tg, sw)
    ref = FO.film_backward(Z, D, p64, tape, go)
    B, Lf, H = c["B"], c["Lf"], 256
    mc_o, film_o = FO.film_core_inputs(Z, p64)
    m = film_model_from_params(c["p"], c["N"], dev)
    mc = t(mc_o.astype(np.float32), dev).requires_grad_(True)
    film = t(film_o.astype(np.float32), dev).requires_grad_(True)
    out = F_.film_decode_core(m.spec, F_.Workspace(), mc, film, t(c["D"], dev), m.core_parameters())
    out.backward(t(go.astype(np.float32), dev))
    torch.cuda.synchronize()
    dfreq = ref["dfreq"].reshape(B, Lf, H)
    dphase = ref["dphase"].reshape(B, Lf, H)
    got = film.grad.cpu().numpy()
    assert O.rel_l2(got[:, :, 0], dfreq[:, 1:]) < TOL_GRAD
    assert O.rel_l2(got[:, :, 1], dphase[:, 1:]) < TOL_GRAD
    # row 4 of d_mc is dL/d(c_b) = sum_p delta_0 = dphase_0
    assert O.rel_l2(mc.grad.cpu().numpy()[:, 4], dphase[:, 0]) < TOL_GRAD


def test_film_dispatch_and_batch_independence(dev):
    torch.manual_seed(0)
    from reni_b200 import RENIAutoDecoderFiLM, get_directions

    m = RENIAutoDecoderFiLM(6, 9, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
    D = get_directions(32).to(dev)
    with torch.no_grad():
        a = m(3, D)
        b = m([1, 3, 4], D.expand(3, -1, -1))
        cc = m(torch.tensor([3, 0], device=dev), D.expand(2, -1, -1))
        d = m(m.Z[[3]], D)
    # (the per-map stage is a handful of torch matmuls whose reduction order depends on the batch size, so the
    # same map decoded in different batches agrees to rounding, not bitwise; a 1e-7 change of freq / phase flips the
    # fp16 rounding of a few activations)
    for x, y in ((a[0], b[1]), (a[0], cc[0])):
        assert O.rel_l2(x.cpu().numpy(), y.cpu().numpy()) < 2e-4
    assert torch.equal(a, d)


def test_film_trainer_steps_match_reference_pattern(dev):
    """RENITrainer on FiLM decoders (autograd over the fused core): one FIT_DECODER step of the auto-decoder vs the
    oracle, a VAD step (KLD in the loss, sampled latents), and a FIT_LATENT step that leaves the decoder untouched."""
    from reni_b200 import RENIAutoDecoderFiLM, RENITrainer, RENIVariationalAutoDecoderFiLM

    torch.manual_seed(0)
    N, W, B = 9, 32, 3
    P = W * W // 2
    m = RENIAutoDecoderFiLM(8, N, "SO2", 256, 3, 64, 2, 3, None, False).to(dev)
    tr = RENITrainer(m, "FIT_DECODER", W, lr=1e-4)
    imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
    idx = torch.tensor([5, 0, 3], device=dev)
    log = tr.training_step((imgs, idx))
    torch.cuda.synchronize()
    f64 = lambda x: x.detach().cpu().numpy().astype(np.float64)  # noqa: E731
    p = FO.FilmParams([f64(l.layer.weight) for l in m.net], [f64(l.layer.bias) for l in m.net], f64(m.final_layer.weight),
                      f64(m.final_layer.bias), [f64(m.mapping_network.network[2 * i].weight) for i in range(3)],
                      [f64(m.mapping_network.network[2 * i].bias) for i in range(3)], "SO2", None)
    Z = f64(m.Z)[[5, 0, 3]]
    D = np.repeat(O.get_directions(W, np.float64), B, 0)
    sw = np.repeat(O.get_sineweight(W, np.float64), B, 0)
    tg = f64(imgs.permute(0, 2, 3, 1).reshape(B, P, 3))
    out_o, tape = FO.film_forward(Z, D, p, tape=True)
    ref = FO.film_backward(Z, D, p, tape, O.loss_grad_wrt_output(out_o, tg, sw))
    assert abs(float(log["loss"]) - O.weighted_mse(out_o, tg, sw)) < 1e-3 * O.weighted_mse(out_o, tg, sw)
    assert O.rel_l2(m.Z.grad[[5, 0, 3]].cpu().numpy(), ref["dZ"]) < TOL_GRAD
    assert float(m.Z.grad[[1, 2, 4, 6, 7]].abs().max()) == 0.0
    assert O.rel_l2(m.net[1].layer.weight.grad.cpu().numpy(), ref["net_dW"][1]) < TOL_GRAD
    assert O.rel_l2(m.mapping_network.network[0].weight.grad.cpu().numpy(), ref["map_dW"][0]) < TOL_GRAD
    assert m.net[1].layer.weight.grad.data_ptr() == tr.flat.views[2].data_ptr()  # grads live in the flat buffer
    tr.optimizer.step()

    mv = RENIVariationalAutoDecoderFiLM(8, N, "SO2", 256, 3, 64, 2, 3, None, False).to(dev)
    trv = RENITrainer(mv, "FIT_DECODER", W, lr=1e-4)
    logv = trv.step((imgs, idx))
    assert set(logv) == {"loss", "mse_loss", "kld_loss"} and mv.mu.grad is not None and mv.log_var.grad is not None
    assert abs(float(logv["loss"]) - float(logv["mse_loss"]) - float(logv["kld_loss"])) < 1e-6

    mf = RENIAutoDecoderFiLM(8, N, "SO2", 256, 3, 64, 2, 3, None, True).to(dev)
    mf.load_state_dict({"model." + k: v for k, v in m.state_dict().items()})
    w_before = mf.net[1].layer.weight.clone()
    trf = RENITrainer(mf, "FIT_LATENT", W, lr=1e-2)
    logf = trf.step((imgs, idx))
    assert set(logf) == {"loss", "mse_loss", "prior_loss", "cosine_loss"}
    assert torch.equal(mf.net[1].layer.weight, w_before) and float(mf.Z.abs().max()) > 0


@pytest.mark.parametrize("task", ["FIT_DECODER", "FIT_LATENT"])
def test_film_trainer_cuda_graph_matches_eager(dev, task):
    """The captured autograd step (mapping network + fused core + loss + backward) replays to the eager results."""
    from reni_b200 import RENIAutoDecoderFiLM, RENITrainer

    torch.manual_seed(1)
    N, W, B = 9, 32, 4
    m = RENIAutoDecoderFiLM(6, N, "SO2", 256, 3, 64, 2, 3, "tanh", task == "FIT_LATENT").to(dev)
    with torch.no_grad():
        m.Z.normal_()
    batches = [(torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1, torch.tensor(ix, device=dev))
               for ix in ([0, 5, 2, 3], [1, 4, 2, 0])]
    eager = RENITrainer(m, task, W)
    ref = []
    for b in batches:
        log = eager.training_step(b)
        ref.append((float(log["loss"]), m.Z.grad.clone(), [p.grad.clone() for p in m.core_parameters() if p.grad is not None]))
    graphed = RENITrainer(m, task, W, cuda_graph=True)
    for rep in range(2):
        for b, (loss, dZ, dps) in zip(batches, ref):
            log = graphed.training_step(b)
            torch.cuda.synchronize()
            assert abs(float(log["loss"]) - loss) < 1e-5 * abs(loss) + 1e-9
            assert O.rel_l2(m.Z.grad.cpu().numpy(), dZ.cpu().numpy()) < 1e-4
            got = [p.grad for p in m.core_parameters() if p.grad is not None]
            assert len(got) == len(dps)
            for a_, b_ in zip(got, dps):
                assert O.rel_l2(a_.cpu().numpy(), b_.cpu().numpy()) < 1e-4


def test_film_full_size_properties(dev):
    """The default FiLM decoder (N=36, 5 FiLM layers, 3x256 mapping network) at the BASELINE configs[1] size
    (32 maps x 64x128) through size-independent properties:
    (1) SO(2) invariance of the radiance; (2) additivity of every decoder gradient over the two halves of the batch and
    per-map independence of dZ; (3) the no-grad decode equals the differentiated forward (to the rounding of the two per-map stages);
    (4) spot check of 2 maps (forward and dZ) against the oracle."""
    torch.manual_seed(7)
    from reni_b200 import RENIAutoDecoderFiLM, RENITrainLoss, get_directions, get_sineweight

    B, N, W = 32, 36, 128
    P = W * W // 2
    m = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
    with torch.no_grad():
        m.mapping_network.network[-1].weight.mul_(2.0)  # move freq / phase away from their init
    D = get_directions(W).to(dev)
    sw = get_sineweight(W).to(dev).expand(B, -1, -1)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = (0.5 * m.Z.detach()).clone()
    with torch.no_grad():
        o1 = m(Z, D.expand(B, -1, -1))
        th = 0.7
        R = torch.tensor([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], device=dev,
                         dtype=torch.float32)
        o2 = m(Z @ R.T, (D @ R.T).expand(B, -1, -1))
    assert float((o1 - o2).norm() / o1.norm()) < TOL_RADIANCE

    def grads(sel):
        for p in m.parameters():
            p.grad = None
        Zs = Z[sel].clone().requires_grad_(True)
        out = m(Zs, D.expand(Zs.shape[0], -1, -1))
        RENITrainLoss()(out, tg[sel], sw[sel]).backward()
        return out.detach(), Zs.grad.clone(), [p.grad.clone() for n, p in m.named_parameters() if n != "Z"]

    o_full, dZ_full, g_full = grads(slice(0, B))
    _, dZ_a, g_a = grads(slice(0, 16))
    _, dZ_b, g_b = grads(slice(16, B))
    # (no-grad decodes of <= 128 latents use the native per-map stage, the differentiated forward the torch one)
    assert float((o_full - o1).norm() / o1.norm()) < 2e-4
    for a, b1, b2 in zip(g_full, g_a, g_b):
        assert float((a - b1 - b2).norm() / a.norm()) < 2e-3
    assert float((dZ_full[:16] - dZ_a).norm() / dZ_a.norm()) < 1e-3
    assert float((dZ_full[16:] - dZ_b).norm() / dZ_b.norm()) < 1e-3

    f64 = lambda x: x.detach().cpu().numpy().astype(np.float64)  # noqa: E731
    nm = len(m.mapping_network.network) // 2 + 1
    p = FO.FilmParams([f64(l.layer.weight) for l in m.net], [f64(l.layer.bias) for l in m.net], f64(m.final_layer.weight),
                      f64(m.final_layer.bias), [f64(m.mapping_network.network[2 * i].weight) for i in range(nm)],
                      [f64(m.mapping_network.network[2 * i].bias) for i in range(nm)], "SO2", None)
    D64, outs, refs = f64(D), [], []
    for b in (3, 20):
        Zb = f64(Z[[b]])
        out_o, tape = FO.film_forward(Zb, D64, p, tape=True)
        ref = FO.film_backward(Zb, D64, p, tape, O.loss_grad_wrt_output(out_o, f64(tg[[b]]), f64(sw[[b]])))
        outs.append(o_full[[b]].cpu().numpy())
        refs.append(out_o)
        assert O.rel_l2(dZ_full[[b]].cpu().numpy(), ref["dZ"]) < TOL_GRAD
    outs, refs = np.concatenate(outs), np.concatenate(refs)
    print("FiLM config-2 radiance over 2 maps: rel-L2", O.rel_l2(outs, refs), "max abs err", np.abs(outs - refs).max())
    assert O.rel_l2(outs, refs) < TOL_RADIANCE and np.abs(outs - refs).max() < 5e-4


@pytest.mark.parametrize("name", FILM_H256)
def test_film_fused_step_matches_reference_golden(dev, name):
    """reni_film_loss_forward_backward (loss and its gradient formed in the kernels) + autograd of the per-map stage
    against the reference: RENITrainLoss with every gradient, RENITestLoss (cosine term) with the latent gradient."""
    from reni_b200 import functional as F_

    c = load_film_case(name)
    g = c["g"]
    m = film_model_from_params(c["p"], c["N"], dev)
    D, tg, sw = (t(c[k], dev) for k in ("D", "target", "sw"))
    Z = t(c["Z"], dev).requires_grad_(True)
    mc, film = m.map_level(Z)
    res = F_.film_loss_forward_backward(m.spec, F_.Workspace(), mc.detach(), film.detach(), D, tg, sw,
                                        m.core_parameters(), need_dw=True)
    torch.autograd.backward([mc, film], [res.d_mc, res.d_film])
    torch.cuda.synchronize()
    assert abs(float(res.loss) - float(g["train_loss_f32"])) < 1e-3 * abs(float(g["train_loss_f32"]))
    assert O.rel_l2(res.out.cpu().numpy(), g["out_f32"]) < TOL_RADIANCE
    assert O.rel_l2(Z.grad.cpu().numpy(), g["train_dZ_f32"]) < TOL_GRAD
    Lf = c["Lf"]
    assert O.rel_l2(sub(m.net[0].layer.weight.grad.cpu().numpy()), g["net_dW0_f32"]) < TOL_GRAD
    for i in range(1, Lf):
        assert O.rel_l2(sub(res.dW[i - 1].cpu().numpy()), g[f"net_dW{i}_f32"]) < TOL_GRAD, i
        assert O.rel_l2(res.db[i - 1].cpu().numpy(), g[f"net_db{i}_f32"]) < TOL_GRAD, i
    assert O.rel_l2(res.dW[Lf - 1].cpu().numpy(), g["final_dW_f32"]) < TOL_GRAD
    assert O.rel_l2(res.db[Lf - 1].cpu().numpy(), g["final_db_f32"]) < TOL_GRAD
    assert O.rel_l2(sub(m.mapping_network.network[0].weight.grad.cpu().numpy()), g["map_dW0_f32"]) < TOL_GRAD

    Z2 = t(c["Z"], dev).requires_grad_(True)
    mc2, film2 = m.map_level(Z2)
    r2 = F_.film_loss_forward_backward(m.spec, F_.Workspace(), mc2.detach(), film2.detach(), D, tg, sw,
                                       m.core_parameters(), beta=c["beta"], use_cosine=True, need_dw=False)
    prior = c["alpha"] * torch.sum(Z2 ** 2)
    torch.autograd.backward([mc2, film2, prior], [r2.d_mc, r2.d_film, torch.ones_like(prior)])
    torch.cuda.synchronize()
    got = [float(r2.loss) + float(prior), float(r2.mse_loss), float(prior), float(r2.cosine_loss)]
    np.testing.assert_allclose(got, g["test_loss_f32"], rtol=1e-3, atol=1e-8)
    assert O.rel_l2(Z2.grad.cpu().numpy(), g["test_dZ_f32"]) < TOL_GRAD
    assert r2.dW is None


@pytest.mark.parametrize("name", ["film_so2_n9_h256", "film_so2_n36_h256", "film_so3_n9_h256_tanh"])
def test_film_native_map_level_matches_torch_and_oracle(dev, name):
    """reni_film_map_forward (one launch: mapping input, mapping network, freq/phase, hoisted first layer) against the
    differentiable torch stage and the fp64 oracle."""
    c = load_film_case(name)
    m = film_model_from_params(c["p"], c["N"], dev)
    Z = t(c["Z"], dev)
    with torch.no_grad():
        mc_t, film_t = m.map_level(Z)
        mc_n, film_n = m._map_level_native(Z)
    torch.cuda.synchronize()
    mc_o, film_o = FO.film_core_inputs(c["Z"].astype(np.float64), c["p"].astype(np.float64))
    assert O.rel_l2(mc_n.cpu().numpy(), mc_o) < 2e-6 and O.rel_l2(film_n.cpu().numpy(), film_o) < 2e-6
    assert O.rel_l2(mc_n.cpu().numpy(), mc_t.cpu().numpy()) < 2e-6
    assert O.rel_l2(film_n.cpu().numpy(), film_t.cpu().numpy()) < 2e-6


def test_graphed_decoder_replays_and_follows_weight_updates(dev):
    """GraphedDecoder: one CUDA-graph replay per decode, for both decoder families; weights are read at replay time."""
    from reni_b200 import GraphedDecoder, RENIAutoDecoder, RENIAutoDecoderFiLM, get_directions

    torch.manual_seed(3)
    D = get_directions(32).to(dev)
    for m, exact in ((RENIAutoDecoder(4, 9, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, True).to(dev), True),
                     (RENIAutoDecoderFiLM(4, 9, "SO2", 256, 5, 256, 3, 3, None, True).to(dev), False)):
        gd = GraphedDecoder(m, 2, D)
        for trial in range(3):
            Z = torch.randn(2, 9, 3, device=dev)
            got = gd(Z).clone()
            with torch.no_grad():
                want = m(Z, D.expand(2, -1, -1))
            if exact:
                assert torch.equal(got, want)
            else:  # (native per-map stage in the graph, torch stage eagerly: equal to rounding)
                assert O.rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 5e-4
            with torch.no_grad():  # an "optimiser step": the next replay must see the new weights
                for p in m.parameters():
                    if p.dim() == 2:
                        p.mul_(1.01)


def test_film_per_map_images_match_in_epilogue_modulation(dev, monkeypatch):
    """The two FiLM cores -- modulation folded into per-map fp16 weight / bias images (reni_film_prepare_maps, the
    default when P % 512 == 0) and modulation applied in the kernel epilogues (RENI_FILM_PERMAP=0, the only path for
    other P) -- on the same inputs: radiance and every gradient agree to well inside the parity tolerances, and a P
    that is not a multiple of 512 is refused by the per-map entry point."""
    import ctypes as C

    from reni_b200 import RENIAutoDecoderFiLM, RENITrainLoss, _lib

    torch.manual_seed(21)
    B, N, P = 5, 9, 1024
    m = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, "tanh", False).to(dev)
    with torch.no_grad():
        m.mapping_network.network[-1].weight.mul_(2.0)
    D = torch.nn.functional.normalize(torch.randn(B, P, 3, device=dev), dim=-1)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    sw = torch.rand(B, P, 1, device=dev).expand(B, P, 3).contiguous()

    def run():
        for p in m.parameters():
            p.grad = None
        Z = m.Z.detach().clone().requires_grad_(True)
        out = m(Z, D)
        RENITrainLoss()(out, tg, sw).backward()
        return out.detach(), Z.grad.clone(), [p.grad.clone() for n, p in m.named_parameters() if n != "Z"]

    o_map, dz_map, g_map = run()
    monkeypatch.setenv("RENI_FILM_PERMAP", "0")
    o_epi, dz_epi, g_epi = run()
    monkeypatch.delenv("RENI_FILM_PERMAP")
    # (two fp16 roundings of the same weights, each within TOL_RADIANCE of the reference: their mutual distance is
    # up to sqrt(2) of it)
    assert float((o_map - o_epi).norm() / o_epi.norm()) < 2 * TOL_RADIANCE
    assert float((dz_map - dz_epi).norm() / dz_epi.norm()) < 0.5 * TOL_GRAD
    for a, b in zip(g_map, g_epi):
        assert float((a - b).norm() / b.norm()) < 0.5 * TOL_GRAD
    lib = _lib.load()
    cfg = _lib.RENIConfig(N, 1, 256, 4, 3, 1, 1, 1.0, 1.0)
    dummy = torch.zeros(1 << 20, device=dev)
    ptrs = (C.c_void_p * 6)(*[dummy.data_ptr()] * 6)
    rc = lib.reni_film_prepare_maps(C.byref(cfg), dummy.data_ptr(), ptrs, ptrs, 2, 640, dummy.data_ptr(), 1 << 22,
                                    _lib.FLAG_FILM, None)
    assert rc != 0


@pytest.mark.parametrize("eq", ["SO2", "SO3"])
def test_native_per_map_stage_matches_torch_stage(dev, eq, monkeypatch):
    """reni_film_map_forward_train / reni_film_map_backward (the per-map stage of a differentiated FiLM decode: mapping
    input, mapping network, freq / phase, hoisted first layer) against the plain torch ops + autograd of
    ``map_level`` on the same inputs: outputs, dZ and every parameter gradient; then with gradient sinks (accumulated
    in place, None returned) as the trainer uses it."""
    from reni_b200 import RENIAutoDecoderFiLM

    torch.manual_seed(31)
    B, N = 7, 9
    m = RENIAutoDecoderFiLM(B, N, eq, 256, 5, 256, 3, 3, None, False).to(dev)
    with torch.no_grad():
        m.mapping_network.network[-1].weight.mul_(3.0)
    params = m._map_params()
    d_mc = torch.randn(B, 5, 256, device=dev)
    d_film = torch.randn(B, 4, 2, 256, device=dev)

    def run(sinks=None):
        Z = m.Z.detach().clone().requires_grad_(True)
        mc, film = m.map_level(Z, grad_sinks=sinks)
        gs = torch.autograd.grad([mc, film], [Z] + params, [d_mc, d_film], allow_unused=True)
        return mc.detach(), film.detach(), gs

    mc_n, film_n, g_n = run()
    monkeypatch.setenv("RENI_FILM_NATIVE_MAP", "0")
    mc_t, film_t, g_t = run()
    monkeypatch.delenv("RENI_FILM_NATIVE_MAP")
    assert float((mc_n - mc_t).norm() / mc_t.norm()) < 1e-5
    assert float((film_n - film_t).norm() / film_t.norm()) < 1e-5
    for a, b in zip(g_n, g_t):
        assert a is not None and float((a - b).norm() / b.norm()) < 1e-4
    sinks = [torch.ones_like(p) for p in params]
    _, _, g_s = run(sinks)
    assert all(g is None for g in g_s[1:])
    assert float((g_s[0] - g_t[0]).norm() / g_t[0].norm()) < 1e-4
    for s_, b in zip(sinks, g_t[1:]):
        assert float((s_ - 1.0 - b).norm() / b.norm()) < 1e-4
    # frozen decoder: only dZ
    for p in m.parameters():
        p.requires_grad_(False)
    Z = m.Z.detach().clone().requires_grad_(True)
    mc, film = m.map_level(Z)
    (dz,) = torch.autograd.grad([mc, film], [Z], [d_mc, d_film])
    assert float((dz - g_t[0]).norm() / g_t[0].norm()) < 1e-4
