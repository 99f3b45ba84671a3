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
