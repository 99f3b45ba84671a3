"""RENITrainer on the GPU: things a training loop does BETWEEN steps (optimiser updates under CUDA-graph replay, batch
shapes that grow the workspace) and the variational auto-decoder step on the library's sample / KLD kernels."""
import numpy as np
import pytest
import torch

from helpers import O  # noqa: F401  (path setup)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    import __graft_entry__ as entry

    entry.build()
    return torch.device("cuda:0")


def test_vad_step_matches_torch_autograd(dev):
    """RENIVariationalAutoDecoder FIT_DECODER step (RENI_module.py:100-101,312-315): reni_vad_sample + fused decoder step
    + reni_vad_backward against the reference's formulation through torch autograd with the SAME noise."""
    from reni_b200 import RENITrainer, RENIVADTrainLoss, RENIVariationalAutoDecoder

    W, B, N = 32, 4, 9
    imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
    idx = torch.tensor([0, 2, 5, 5], device=dev)  # a repeated row: gradients accumulate like index_add_
    v = RENIVariationalAutoDecoder(8, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    with torch.no_grad():
        v.log_var.add_(4.0)  # std ~ 0.6: the sampled term matters
    tr = RENITrainer(v, "FIT_DECODER", W, lr=1e-4, kld_weighting=1e-2)
    torch.manual_seed(77)
    log = tr.training_step((imgs, idx))
    torch.cuda.synchronize()
    got_mu, got_lv = v.mu.grad.clone(), v.log_var.grad.clone()
    got_w = [p.grad.clone() for p in v.net.parameters()]
    # reference formulation, same generator state -> same eps
    torch.manual_seed(77)
    for p in v.parameters():
        p.grad = None
    eps = torch.randn(B, N, 3, device=dev)
    mu, lv = v.mu[idx], v.log_var[idx]
    Z = mu + eps * torch.exp(0.5 * lv)
    out = v(Z, tr.directions.expand(B, -1, -1))
    t = imgs.permute(0, 2, 3, 1).reshape(B, -1, 3)
    loss, mse, kld = RENIVADTrainLoss(beta=1e-2, Z_dims=3 * N)(out, t, tr.sineweight.expand(B, -1, -1), mu, lv)
    loss.backward()
    torch.cuda.synchronize()
    assert set(log) == {"loss", "mse_loss", "kld_loss"}
    assert abs(float(log["kld_loss"]) - float(kld)) <= 1e-5 * abs(float(kld))
    assert abs(float(log["loss"]) - float(loss)) <= 1e-4 * abs(float(loss))
    assert float((got_mu - v.mu.grad).norm() / v.mu.grad.norm()) < 2e-3
    assert float((got_lv - v.log_var.grad).norm() / v.log_var.grad.norm()) < 2e-3
    for a, p in zip(got_w, v.net.parameters()):
        assert float((a - p.grad).norm() / p.grad.norm()) < 2e-3
    assert float(got_mu[[1, 3, 4, 6, 7]].abs().max()) == 0.0


def test_vad_step_under_cuda_graph(dev):
    """The VAD step is graph-capturable (noise drawn by torch.randn inside the capture): replays draw fresh noise, the KLD
    value (noise-free) matches torch, and training makes progress."""
    from reni_b200 import KLD, RENITrainer, RENIVariationalAutoDecoder

    torch.manual_seed(3)
    W, B, N = 32, 4, 9
    imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
    idx = torch.tensor([1, 2, 5, 7], device=dev)
    v = RENIVariationalAutoDecoder(8, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    tr = RENITrainer(v, "FIT_DECODER", W, lr=1e-4, kld_weighting=1e-3, cuda_graph=True)
    first = tr.step((imgs, idx))
    torch.cuda.synchronize()
    k0 = float(first["kld_loss"])
    assert tr._graphs, "the VAD step must have been captured"
    outs = []
    for _ in range(3):
        tr.training_step((imgs, idx))
        torch.cuda.synchronize()
        outs.append(tr.last_output.clone())
    assert not torch.equal(outs[0], outs[1])  # fresh noise per replay
    want = 1e-3 * float(KLD(v.mu[idx], v.log_var[idx], Z_dims=3 * N))
    lg = tr.training_step((imgs, idx))
    assert abs(float(lg["kld_loss"]) - want) <= 1e-4 * abs(want) and k0 > 0
    l0 = float(lg["mse_loss"])
    for _ in range(40):
        lg = tr.step((imgs, idx))
    assert float(lg["mse_loss"]) < l0


@pytest.mark.parametrize("W", [16, 32])  # P = 128 (shared weight images only) and P = 512 (per-map weight images)
@pytest.mark.parametrize("film", [False, True])
def test_graphed_trainer_follows_optimizer_steps(dev, W, film):
    """trainer.step() = training_step + Adam, repeated: a captured graph must rebuild the fp16 weight images from the
    UPDATED parameters in every replay.  Eager and graphed runs from the same initial state give the same loss sequence,
    and the loss moves."""
    from reni_b200 import RENIAutoDecoder, RENIAutoDecoderFiLM, RENITrainer

    torch.manual_seed(5)
    B, N = 4, 9
    mk = (lambda: RENIAutoDecoderFiLM(8, N, "SO2", 256, 5, 256, 3, 3, None, False)) if film else \
        (lambda: RENIAutoDecoder(8, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False))
    a = mk().to(dev)
    b = mk().to(dev)
    b.load_state_dict({"model." + k: v for k, v in a.state_dict().items()})  # (Lightning-prefixed keys, RENI.py:190-203)
    imgs = torch.rand(B, 3, W // 2, W, device=dev) * 2 - 1
    idx = torch.tensor([0, 3, 4, 6], device=dev)
    te = RENITrainer(a, "FIT_DECODER", W, lr=3e-4)
    tg = RENITrainer(b, "FIT_DECODER", W, lr=3e-4, cuda_graph=True)
    le, lg = [], []
    for _ in range(8):
        le.append(float(te.step((imgs, idx))["loss"]))
        lg.append(float(tg.step((imgs, idx))["loss"]))
    torch.cuda.synchronize()
    assert le[-1] < le[0] and abs(le[-1] - le[0]) > 1e-3 * abs(le[0])
    for x, y in zip(le, lg):
        assert abs(x - y) <= 2e-3 * abs(x), (le, lg)
    for p, q in zip(a.parameters(), b.parameters()):
        assert float((p - q).abs().max()) <= 1e-4 + 1e-2 * float((p.abs().max()))


def test_graphs_keep_their_own_workspace(dev):
    """A graph captured for a small batch must survive a later, larger batch shape (which needs a bigger workspace) and
    an eager decode in between: every captured graph owns its workspace."""
    from reni_b200 import GraphedDecoder, RENIAutoDecoder, RENITrainer, get_directions

    torch.manual_seed(9)
    W, N = 32, 9
    m = RENIAutoDecoder(16, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    tg = RENITrainer(m, "FIT_DECODER", W, lr=1e-4, cuda_graph=True)
    te = RENITrainer(m, "FIT_DECODER", W, lr=1e-4)
    small = (torch.rand(2, 3, W // 2, W, device=dev) * 2 - 1, torch.tensor([1, 4], device=dev))
    big = (torch.rand(12, 3, W // 2, W, device=dev) * 2 - 1, torch.arange(12, device=dev))

    def check(batch):
        lg = tg.training_step(batch)
        torch.cuda.synchronize()
        gz, gw, gl = m.Z.grad.clone(), [p.grad.clone() for p in m.net.parameters()], float(lg["loss"])
        le = te.training_step(batch)
        torch.cuda.synchronize()
        assert abs(gl - float(le["loss"])) <= 1e-6 * abs(gl)
        assert float((gz - m.Z.grad).abs().max()) <= 1e-5 * float(m.Z.grad.abs().max()) + 1e-12
        for x, p in zip(gw, m.net.parameters()):
            assert float((x - p.grad).abs().max()) <= 2e-4 * float(p.grad.abs().max()) + 1e-12

    check(small)
    check(big)    # grows every shared workspace
    check(small)  # the first graph replays on memory it still owns
    D = get_directions(W).to(dev)
    gd = GraphedDecoder(m, 2, D)
    Z2 = torch.randn(2, N, 3, device=dev)
    ref = m(Z2, D.expand(2, -1, -1)).clone()
    with torch.no_grad():
        m(torch.randn(12, N, 3, device=dev), D.expand(12, -1, -1))  # a larger eager decode in between
    assert torch.equal(gd(Z2), ref)


@pytest.mark.parametrize("masked,cosine", [(False, False), (True, True)])
@pytest.mark.parametrize("tile_major", [True, False])
def test_analytic_grid_and_bitmask_match_arrays(dev, masked, cosine, tile_major):
    """RENI_FLAG_GRID_DIRECTIONS / RENI_FLAG_GRID_SINEWEIGHT: directions and sine weights computed in the kernels from
    the pixel index (utils.py:46-78) and the mask as one bit per pixel (RENI_module.py:92-94) give the step computed from
    the get_directions / get_sineweight * mask arrays; both backward schedules; the trainer option too."""
    from reni_b200 import RENIAutoDecoder, RENITrainer, get_directions, get_sineweight, pack_mask_bits, rectangle_mask
    from reni_b200 import functional as F_

    torch.manual_seed(4)
    W, B, N = 64, 3, 9
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
    mask = rectangle_mask(W, 5, 23, 20, 41).to(dev) if masked else None
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = m.Z.detach()
    kw = dict(alpha=1e-3 if cosine else 0.0, beta=0.3 if cosine else 0.0, use_cosine=cosine, need_dw=True,
              tile_major_bwd=tile_major)
    a = F_.loss_forward_backward(m.spec, F_.Workspace(), Z, D, tg, sw if mask is None else sw * mask,
                                 m.decoder_weights(), m.decoder_biases(), **kw)
    b = F_.loss_forward_backward(m.spec, F_.Workspace(), Z, None, tg, None, m.decoder_weights(), m.decoder_biases(),
                                 mask_bits=pack_mask_bits(mask) if masked else None, **kw)
    torch.cuda.synchronize()
    # the grid in closed form (sincospi) and the arrays (torch.sin of a rounded pi * x) differ in the last bits, which
    # flips the fp16 rounding of an activation here and there: compare at that level
    assert float((a.out - b.out).norm() / a.out.norm()) < 1e-4 and float((a.out - b.out).abs().max()) < 2e-4
    assert abs(float(a.loss) - float(b.loss)) < 1e-4 * abs(float(a.loss))
    assert float((a.dZ - b.dZ).norm() / a.dZ.norm()) < 2e-3
    for x, y in zip(a.dW + a.db, b.dW + b.db):
        assert float((x - y).norm() / x.norm()) < 2e-3
    imgs = tg.reshape(B, W // 2, W, 3).permute(0, 3, 1, 2).contiguous()
    idx = torch.arange(B, device=dev)
    t0 = RENITrainer(m, "FIT_LATENT" if cosine else "FIT_DECODER", W, mask=mask, analytic_grid=False)
    t1 = RENITrainer(m, "FIT_LATENT" if cosine else "FIT_DECODER", W, mask=mask, analytic_grid=True)
    l0 = t0.training_step((imgs, idx))
    g0 = m.Z.grad.clone()
    l1 = t1.training_step((imgs, idx))
    torch.cuda.synchronize()
    assert abs(float(l0["loss"]) - float(l1["loss"])) < 1e-4 * abs(float(l0["loss"]))
    assert float((g0 - m.Z.grad).norm() / g0.norm()) < 2e-3


def test_large_batches_are_walked_in_chunks_of_maps(dev, monkeypatch):
    """A batch whose stash exceeds the workspace budget is processed in chunks of maps through one workspace (maps are
    independent units, gradients accumulate): same loss, outputs and gradients as the single call."""
    from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight
    from reni_b200 import functional as F_

    torch.manual_seed(2)
    W, B, N = 32, 11, 9
    P = W * W // 2
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    D, sw = get_directions(W).to(dev), get_sineweight(W).to(dev)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    kw = dict(alpha=1e-3, beta=0.2, use_cosine=True, need_dw=True)
    a = F_.loss_forward_backward(m.spec, F_.Workspace(), m.Z.detach(), D, tg, sw, m.decoder_weights(), m.decoder_biases(), **kw)
    ws = F_.Workspace()
    flags = 1 | 2 | 4  # SAVE_FOR_BACKWARD | NEED_DW | LOSS
    room = F_.workspace_bytes(m.spec.c_config(), 4, P, flags)  # room for four maps of this size
    monkeypatch.setenv("RENI_MAX_WORKSPACE_GB", repr(room / 2 ** 30))
    b = F_.loss_forward_backward(m.spec, ws, m.Z.detach(), D, tg, sw, m.decoder_weights(), m.decoder_biases(), **kw)
    torch.cuda.synchronize()
    assert ws.nbytes <= room < F_.workspace_bytes(m.spec.c_config(), B, P, flags)
    assert torch.equal(a.out, b.out)
    for x, y in zip((a.loss, a.mse_loss, a.prior_loss, a.cosine_loss), (b.loss, b.mse_loss, b.prior_loss, b.cosine_loss)):
        assert abs(float(x) - float(y)) <= 1e-5 * abs(float(x)) + 1e-9
    assert float((a.dZ - b.dZ).abs().max()) <= 1e-5 * float(a.dZ.abs().max())
    for x, y in zip(a.dW + a.db, b.dW + b.db):
        assert float((x - y).norm() / x.norm()) < 1e-4
