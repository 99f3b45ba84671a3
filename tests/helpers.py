"""Shared test helpers: golden case loading and conversions between the oracle's DecoderParams and the
reni_b200 modules."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import reni_oracle as O  # noqa: E402
from make_golden import CASES, golden_inputs, sub_dw  # noqa: E402,F401

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity bars of BASELINE.json's north_star, written once here
TOL_RADIANCE = 1e-3   # radiance, relative (rel-L2) vs the reference fp32 decoder
TOL_GRAD = 1e-2       # weight and latent gradients, relative (rel-L2)
TOL_RADIANCE_MAX = 3e-3  # max-abs/max-abs is also reported; fp16 operands put it at ~1.2e-3


def load_case(name):
    seed, B, P, N, H, L, out_f, eq, last_lin, act, grid, alpha, beta, full = CASES[name]
    p, Z, D, target, sw, mask = golden_inputs(seed, B, P, N, H, L, out_f, eq, grid_sidelen=grid)
    p.last_layer_linear = last_lin
    p.output_activation = act
    if name.endswith("masked"):
        sw = sw * mask
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return dict(p=p, Z=Z, D=D, target=target, sw=sw, g=g, alpha=alpha, beta=beta, full=full, N=N, B=B, P=P, L=L)


def model_from_params(p, N, device, dataset_size=4, fixed=False, cls=None):
    import torch

    from reni_b200 import RENIAutoDecoder

    cls = cls or RENIAutoDecoder
    H = p.weights[1].shape[0]
    L = len(p.weights) - 2
    m = cls(dataset_size, N, p.equivariance, H, L, p.weights[-1].shape[0], p.last_layer_linear,
            p.output_activation, p.first_omega_0, p.hidden_omega_0, fixed)
    with torch.no_grad():
        for w, b, tw, tb in zip(p.weights, p.biases, m.decoder_weights(), m.decoder_biases()):
            tw.copy_(torch.from_numpy(np.asarray(w, dtype=np.float32)))
            tb.copy_(torch.from_numpy(np.asarray(b, dtype=np.float32)))
    return m.to(device)


def params_from_model(m, dtype=np.float64):
    return O.DecoderParams([w.detach().cpu().numpy().astype(dtype) for w in m.decoder_weights()],
                           [b.detach().cpu().numpy().astype(dtype) for b in m.decoder_biases()],
                           m.first_omega_0, m.hidden_omega_0, bool(m.last_layer_linear), m.output_activation,
                           m.equivariance)
