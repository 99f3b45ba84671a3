"""FusedAdam (reni_adam_step) against torch.optim.Adam(params, lr) -- the optimiser the reference constructs
(src/lightning/RENI_module.py:185-192) -- over several steps, with a per-epoch ExponentialLR (:212-214), parameters of
odd sizes spanning more than one launch (> 24 segments), and a dense latent table whose untouched rows still move."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__ as entry

    entry.build()
    return torch.device("cuda:0")


def test_fused_adam_matches_torch_adam(dev):
    from reni_b200 import FusedAdam

    torch.manual_seed(0)
    shapes = [(256, 1370), (256,), (256, 256), (3, 256), (3,), (64, 36, 3), (1,), (7, 5)] + [(33,)] * 22
    pa = [torch.randn(s, device=dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = FusedAdam(pa, lr=1e-2)
    ob = torch.optim.Adam(pb, lr=1e-2)
    sa = torch.optim.lr_scheduler.ExponentialLR(oa, gamma=0.9)
    sb = torch.optim.lr_scheduler.ExponentialLR(ob, gamma=0.9)
    for step in range(12):
        for i, (a, b) in enumerate(zip(pa, pb)):
            g = torch.randn_like(a) * (10.0 ** ((i % 5) - 3))
            if a.dim() == 3:
                g[::2] = 0  # latent table: rows outside the batch have zero gradient but keep their momentum
            a.grad = g.clone()
            b.grad = g.clone()
        oa.step()
        ob.step()
        if step % 4 == 3:
            sa.step()
            sb.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=2e-5, atol=1e-7)
    for a, b in zip(pa, pb):
        np.testing.assert_allclose(oa.state[a]["exp_avg_sq"].cpu().numpy(), ob.state[b]["exp_avg_sq"].cpu().numpy(),
                                   rtol=1e-5, atol=1e-12)
    assert int(oa.param_groups[0]["_step"]) == 12


def test_trainer_uses_fused_adam_and_moves_parameters(dev):
    from reni_b200 import FusedAdam, RENIAutoDecoder, RENITrainer

    torch.manual_seed(0)
    m = RENIAutoDecoder(8, 9, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    ref = RENIAutoDecoder(8, 9, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    ref.load_state_dict({"model." + k: v for k, v in m.state_dict().items()})
    tr = RENITrainer(m, "FIT_DECODER", 32, lr=1e-3)
    tr2 = RENITrainer(ref, "FIT_DECODER", 32, lr=1e-3)
    tr2.optimizer = torch.optim.Adam(list(ref.parameters()), lr=1e-3)
    assert isinstance(tr.optimizer, FusedAdam)
    imgs = torch.rand(3, 3, 16, 32, device=dev) * 2 - 1
    idx = torch.tensor([5, 0, 3], device=dev)
    for _ in range(3):
        tr.step((imgs, idx))
        tr2.step((imgs, idx))
    torch.cuda.synchronize()
    for a, b in zip(m.parameters(), ref.parameters()):
        # (Adam's first steps are sign-like: a gradient whose rounding differs in the last bit can move an element by
        # up to 2 lr, so compare in aggregate)
        d = (a - b).abs()
        assert float(d.mean()) < 1e-5 and float(d.max()) < 7e-3
