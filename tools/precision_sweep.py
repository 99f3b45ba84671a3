"""Radiance / gradient error of the fp16-operand kernels vs the fp64 oracle over several random-init seeds
(reference constructor init), to characterise the 1e-3 / 1e-2 tolerances.  Run under gpurun."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import reni_oracle as O
from helpers import params_from_model
from reni_b200 import RENIAutoDecoder, get_directions, get_sineweight
from reni_b200 import functional as F_
dev = torch.device("cuda:0")
W, B, N = 64, 4, 36
P = W * W // 2
D = get_directions(W).to(dev); sw = get_sineweight(W).to(dev)
D64 = np.repeat(D.cpu().numpy().astype(np.float64), B, 0); sw64 = np.repeat(sw.cpu().numpy().astype(np.float64), B, 0)
for seed in range(8):
    torch.manual_seed(seed)
    m = RENIAutoDecoder(B, N, "SO2", 256, 5, 3, True, "tanh", 30.0, 30.0, False).to(dev)
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    r = F_.loss_forward_backward(m.spec, F_.Workspace(), m.Z.detach(), D, tg, sw, m.decoder_weights(), m.decoder_biases())
    ref = O.step_fit_decoder(m.Z.detach().cpu().numpy().astype(np.float64), D64, tg.cpu().numpy().astype(np.float64), sw64, params_from_model(m))
    o = r.out.cpu().numpy()
    edw = max(O.rel_l2(a.cpu().numpy(), b) for a, b in zip(r.dW, ref["dW"]))
    print(f"seed {seed}: radiance rel_l2={O.rel_l2(o, ref['out']):.2e} rel_max={O.rel_max(o, ref['out']):.2e} "
          f"abs_max_err={np.abs(o-ref['out']).max():.2e} out_rms={np.sqrt((ref['out']**2).mean()):.3f} "
          f"dZ={O.rel_l2(r.dZ.cpu().numpy(), ref['dZ']):.2e} max dW={edw:.2e}")

# ---- FiLM decoder (default config: 5 FiLM layers, 3 x 256 mapping network), fused core step + autograd of the per-map stage
import reni_film_oracle as FO
from reni_b200 import RENIAutoDecoderFiLM
f64 = lambda x: x.detach().cpu().numpy().astype(np.float64)  # noqa: E731
for seed in range(8):
    torch.manual_seed(100 + seed)
    m = RENIAutoDecoderFiLM(B, N, "SO2", 256, 5, 256, 3, 3, None, False).to(dev)
    with torch.no_grad():
        m.mapping_network.network[-1].weight.mul_(2.0)  # spread freq / phase as a trained mapping network does
    tg = torch.rand(B, P, 3, device=dev) * 2 - 1
    Z = (0.5 * m.Z.detach()).clone().requires_grad_(True)
    mc, film = m.map_level(Z)
    r = F_.film_loss_forward_backward(m.spec, F_.Workspace(), mc.detach(), film.detach(), D, tg, sw, m.core_parameters())
    torch.autograd.backward([mc, film], [r.d_mc, r.d_film])
    nm = len(m.mapping_network.network) // 2 + 1
    p = FO.FilmParams([f64(l.layer.weight) for l in m.net], [f64(l.layer.bias) for l in m.net], f64(m.final_layer.weight),
                      f64(m.final_layer.bias), [f64(m.mapping_network.network[2 * i].weight) for i in range(nm)],
                      [f64(m.mapping_network.network[2 * i].bias) for i in range(nm)], "SO2", None)
    out_o, tape = FO.film_forward(f64(Z), D64, p, tape=True)
    ref = FO.film_backward(f64(Z), D64, p, tape, O.loss_grad_wrt_output(out_o, f64(tg), sw64))
    o = r.out.cpu().numpy()
    edw = max(O.rel_l2(a.cpu().numpy(), b) for a, b in zip(r.dW[:-1], ref["net_dW"][1:]))
    emap = max(O.rel_l2(m.mapping_network.network[2 * i].weight.grad.cpu().numpy(), ref["map_dW"][i]) for i in range(nm))
    print(f"film seed {seed}: radiance rel_l2={O.rel_l2(o, out_o):.2e} rel_max={O.rel_max(o, out_o):.2e} "
          f"abs_max_err={np.abs(o-out_o).max():.2e} out_rms={np.sqrt((out_o**2).mean()):.3f} "
          f"dZ={O.rel_l2(Z.grad.cpu().numpy(), ref['dZ']):.2e} max net dW={edw:.2e} max mapping dW={emap:.2e}")
