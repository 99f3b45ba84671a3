"""CPU baseline port of the reference decoder step in plain PyTorch (fp32, eager, autograd) --
TEST / BENCHMARK INFRASTRUCTURE ONLY (used by bench.py's cpu_baseline / --impl reference legs and
by tests; never by the product).

The reference's own CPU implementation IS eager PyTorch, but it cannot travel to the GPU box
(/root/reference is absent there), so this restates the op sequence it executes on that path:
the (B, P, 2N+N^2+2) encoding via bmm / repeat / cat (src/models/RENI.py:31-53), six
sin(omega * addmm) layers, Linear + tanh (RENI.py:86-87,132-178), the sine-weighted MSE
(src/utils/loss_functions.py:6-13) and loss.backward() (src/lightning/RENI_module.py:105-118).
tests/test_oracle_golden.py pins it against the fixtures generated from the real reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def so2_encoding(Z: torch.Tensor, D: torch.Tensor) -> torch.Tensor:
    """RENI.py:31-53 -- materialises the full encoding exactly like the reference does."""
    P = D.shape[1]
    Z_xz = torch.stack((Z[..., 0], Z[..., 2]), -1)
    D_xz = torch.stack((D[..., 0], D[..., 2]), -1)
    G = torch.bmm(Z_xz, Z_xz.transpose(1, 2))
    Z_inv = G.flatten(start_dim=1).unsqueeze(1).repeat(1, P, 1)
    ip = torch.bmm(D_xz, Z_xz.transpose(1, 2))
    dn = torch.sqrt(D[..., 0] ** 2 + D[..., 2] ** 2).unsqueeze(2)
    Zy = Z[..., 1].unsqueeze(1).repeat(1, P, 1)
    return torch.cat((ip, Z_inv, dn, Zy, D[..., 1].unsqueeze(2)), 2)


def decoder(Z, D, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], omega: float = 30.0,
            output_activation: Optional[str] = "tanh") -> torch.Tensor:
    h = so2_encoding(Z, D)
    n = len(weights)
    for i in range(n - 1):
        h = torch.sin(omega * torch.nn.functional.linear(h, weights[i], biases[i]))  # RENI.py:87
    h = torch.nn.functional.linear(h, weights[-1], biases[-1])
    return torch.tanh(h) if output_activation == "tanh" else h


def weighted_mse(o, t, sw):
    return (((o - t) ** 2) * sw).view(o.shape[0], -1).mean(1).sum(0)


def training_step(Z, D, target, sw, weights: List[torch.Tensor], biases: List[torch.Tensor]):
    """forward + RENITrainLoss + backward for every parameter and the latents; returns (loss, out, grads)."""
    params = [Z] + list(weights) + list(biases)
    for p in params:
        p.requires_grad_(True)
        p.grad = None
    out = decoder(Z, D, weights, biases)
    loss = weighted_mse(out, target, sw)
    loss.backward()
    return loss.detach(), out.detach(), [p.grad for p in params]
