"""CPU oracle for the FiLM-conditioned RENI decoder -- TEST INFRASTRUCTURE ONLY.

numpy restatement of ``RENIAutoDecoderFiLM`` (JADGardner/RENI ``src/models/RENI.py:407-678``): the
invariant split into a SIREN input and a mapping-network input (``:405-452``), the mapping network
(``:481-512``), the FiLM layers ``sin(freq * (W x + b) + phase)`` (``:515-524``) and
``forward_with_frequencies_phase_shifts`` (``:666-678``), plus a hand-derived reverse pass (what autograd
does for the reference).  Only ``tests/`` may import it.  Pinned against outputs of the reference itself:
``oracle/make_golden_film.py`` -> ``tests/golden/film_*.npz`` -> ``tests/test_film_oracle_golden.py``.

Like the reference, the direct formulation evaluates the mapping network on a per-pixel replicated
input; because that input is constant per map the oracle evaluates it once per map (``per_map=True``,
the hoisted form the CUDA path uses) or per pixel (``per_map=False``, op for op as the reference) --
the golden test proves both equal the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class FilmParams:
    """Weights of ``RENIAutoDecoderFiLM``: ``net[i].layer`` (FiLM layers), ``final_layer`` and
    ``mapping_network.network`` (Linear / LeakyReLU(0.2) stack, last Linear without activation)."""

    net_w: List[np.ndarray]   # [ (H, in0), (H, H) x (Lf - 1) ]
    net_b: List[np.ndarray]
    final_w: np.ndarray       # (out, H)
    final_b: np.ndarray
    map_w: List[np.ndarray]   # mapping network linears, (out, in) each
    map_b: List[np.ndarray]
    equivariance: str = "SO2"
    output_activation: Optional[str] = None  # None | "tanh" | "exp"

    def astype(self, dt) -> "FilmParams":
        c = lambda xs: [x.astype(dt) for x in xs]  # noqa: E731
        return FilmParams(c(self.net_w), c(self.net_b), self.final_w.astype(dt), self.final_b.astype(dt),
                          c(self.map_w), c(self.map_b), self.equivariance, self.output_activation)


def film_in_features(ndims: int, equivariance: str):
    """(SIREN in_features, mapping-network in_features) -- RENI.py:556-567."""
    if equivariance == "SO2":
        return 2 + ndims, ndims * ndims + ndims
    if equivariance == "SO3":
        return ndims, ndims * ndims
    raise ValueError("equivariance 'None' + FiLM is broken in the reference (first layer expects 3N inputs, gets N)")


def film_init(rng: np.random.Generator, ndims: int, equivariance: str = "SO2", hidden: int = 256,
              siren_layers: int = 5, map_features: int = 256, map_layers: int = 3, out_features: int = 3,
              output_activation: Optional[str] = None, dtype=np.float32) -> FilmParams:
    """Same distributions as the reference initialisers (RENI.py:455-478,500-502,584-586); not torch's stream."""
    nin, mn_in = film_in_features(ndims, equivariance)
    net_w, net_b = [], []
    fan = nin
    for i in range(siren_layers):
        lim = (1.0 / fan) if i == 0 else np.sqrt(6 / fan) / 25
        net_w.append(rng.uniform(-lim, lim, (hidden, fan)))
        net_b.append(rng.uniform(-1 / np.sqrt(fan), 1 / np.sqrt(fan), hidden))
        fan = hidden
    lim = np.sqrt(6 / hidden) / 25
    final_w = rng.uniform(-lim, lim, (out_features, hidden))
    final_b = rng.uniform(-1 / np.sqrt(hidden), 1 / np.sqrt(hidden), out_features)
    map_w, map_b = [], []
    fan = mn_in
    gain = np.sqrt(2.0 / (1 + 0.2**2))
    for _ in range(map_layers):
        map_w.append(rng.standard_normal((map_features, fan)) * gain / np.sqrt(fan))
        map_b.append(rng.uniform(-1 / np.sqrt(fan), 1 / np.sqrt(fan), map_features))
        fan = map_features
    n_out = siren_layers * hidden * 2
    map_w.append(0.25 * rng.standard_normal((n_out, fan)) * gain / np.sqrt(fan))
    map_b.append(rng.uniform(-1 / np.sqrt(fan), 1 / np.sqrt(fan), n_out))
    p = FilmParams(net_w, net_b, final_w, final_b, map_w, map_b, equivariance, output_activation)
    return p.astype(dtype)


# --------------------------------------------------------------------------------------
# invariant inputs (RENI.py:405-452), per map for the mapping network
# --------------------------------------------------------------------------------------


def film_inputs(Z: np.ndarray, D: np.ndarray, equivariance: str):
    """-> (Siren_Input (B,P,in), Mapping_Input per map (B, mn_in))."""
    B, N, _ = Z.shape
    if equivariance == "SO2":
        Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)
        D_xz = np.stack((D[:, :, 0], D[:, :, 2]), -1)
        G = Z_xz @ np.transpose(Z_xz, (0, 2, 1))
        innerprod = D_xz @ np.transpose(Z_xz, (0, 2, 1))
        d_norm = np.sqrt(D[:, :, 0] ** 2 + D[:, :, 2] ** 2)[:, :, None]
        d_y = D[:, :, 1][:, :, None]
        siren_in = np.concatenate((d_norm, d_y, innerprod), 2)            # RENI.py:434
        map_in = np.concatenate((G.reshape(B, N * N), Z[:, :, 1]), 1)     # RENI.py:435
        return siren_in, map_in
    if equivariance == "SO3":
        G = Z @ np.transpose(Z, (0, 2, 1))
        return D @ np.transpose(Z, (0, 2, 1)), G.reshape(B, N * N)        # RENI.py:407-415
    raise ValueError(equivariance)


def mapping_forward(x: np.ndarray, p: FilmParams, tape: bool = False):
    """CustomMappingNetwork.forward (RENI.py:504-512) -> raw (.., 2*Lf*H) [freq | phase]."""
    pre = []
    h = x
    n = len(p.map_w)
    for i in range(n):
        y = h @ p.map_w[i].T + p.map_b[i]
        pre.append((h, y))
        h = np.where(y > 0, y, 0.2 * y) if i < n - 1 else y
    return (h, pre) if tape else h


def mapping_backward(g: np.ndarray, p: FilmParams, pre):
    """Reverse of ``mapping_forward`` -> (dW list, db list, dx)."""
    n = len(p.map_w)
    dWs, dbs = [None] * n, [None] * n
    for i in reversed(range(n)):
        h_in, y = pre[i]
        if i < n - 1:
            g = g * np.where(y > 0, 1.0, 0.2).astype(g.dtype)
        dWs[i] = g.reshape(-1, g.shape[-1]).T @ h_in.reshape(-1, h_in.shape[-1])
        dbs[i] = g.reshape(-1, g.shape[-1]).sum(0)
        g = g @ p.map_w[i]
    return dWs, dbs, g


def film_forward(Z: np.ndarray, D: np.ndarray, p: FilmParams, tape: bool = False):
    """RENIAutoDecoderFiLM.forward on latent codes (RENI.py:653-678)."""
    dt = Z.dtype
    x, map_in = film_inputs(Z, D, p.equivariance)
    raw, mtape = mapping_forward(map_in, p, tape=True)
    Lf = len(p.net_w)
    H = p.net_w[0].shape[0]
    freq = raw[:, : Lf * H] * dt.type(15) + dt.type(30)   # RENI.py:667
    phase = raw[:, Lf * H:]
    h = x
    us, pres, acts = [], [], [x]
    for i in range(Lf):
        u = h @ p.net_w[i].T + p.net_b[i]                                  # FiLMLayer.layer
        a = freq[:, None, i * H:(i + 1) * H] * u + phase[:, None, i * H:(i + 1) * H]
        h = np.sin(a)                                                      # RENI.py:524
        us.append(u)
        pres.append(a)
        acts.append(h)
    y = h @ p.final_w.T + p.final_b
    if p.output_activation == "tanh":
        o = np.tanh(y)
    elif p.output_activation == "exp":
        o = np.exp(y)
    else:
        o = y
    if tape:
        return o, dict(x=x, map_in=map_in, mtape=mtape, freq=freq, phase=phase, us=us, pres=pres, acts=acts, y=y, out=o)
    return o


def film_core_inputs(Z: np.ndarray, p: FilmParams):
    """Hoisted per-map operands the CUDA core consumes: mc (B,5,H) with a0 = [f|1] . mc and
    film (B, Lf-1, 2, H) = (freq_l, phase_l) of the hidden FiLM layers."""
    B, N, _ = Z.shape
    Lf, H = len(p.net_w), p.net_w[0].shape[0]
    _, map_in = film_inputs(Z, np.zeros((B, 1, 3), Z.dtype), p.equivariance)
    raw = mapping_forward(map_in, p)
    freq = (raw[:, : Lf * H] * 15 + 30).reshape(B, Lf, H)
    phase = raw[:, Lf * H:].reshape(B, Lf, H)
    W0, b0 = p.net_w[0], p.net_b[0]
    M = np.zeros((B, 4, H), Z.dtype)
    if p.equivariance == "SO2":
        M[:, 0] = Z[:, :, 0] @ W0[:, 2:].T
        M[:, 1] = Z[:, :, 2] @ W0[:, 2:].T
        M[:, 2] = W0[:, 0]
        M[:, 3] = W0[:, 1]
    else:
        for i in range(3):
            M[:, i] = Z[:, :, i] @ W0.T
    mc = np.concatenate((M * freq[:, :1], (freq[:, 0] * b0 + phase[:, 0])[:, None]), 1)
    film = np.stack((freq[:, 1:], phase[:, 1:]), 2)
    return mc, film


def film_backward(Z: np.ndarray, D: np.ndarray, p: FilmParams, t: dict, grad_out: np.ndarray):
    """Reverse pass of ``film_forward`` -> dict(net_dW, net_db, final_dW, final_db, map_dW, map_db, dZ)."""
    B, N, _ = Z.shape
    Lf, H = len(p.net_w), p.net_w[0].shape[0]
    g = grad_out
    if p.output_activation == "tanh":
        g = g * (1 - t["out"] ** 2)
    elif p.output_activation == "exp":
        g = g * t["out"]
    g2 = g.reshape(-1, g.shape[-1])
    final_dW = g2.T @ t["acts"][Lf].reshape(-1, H)
    final_db = g2.sum(0)
    g = g @ p.final_w
    dfreq = np.zeros_like(t["freq"])
    dphase = np.zeros_like(t["phase"])
    net_dW, net_db = [None] * Lf, [None] * Lf
    for i in reversed(range(Lf)):
        da = g * np.cos(t["pres"][i])                       # dL/da_i
        dphase[:, i * H:(i + 1) * H] = da.sum(1)
        dfreq[:, i * H:(i + 1) * H] = (da * t["us"][i]).sum(1)
        du = da * t["freq"][:, None, i * H:(i + 1) * H]
        h_in = t["acts"][i]
        net_dW[i] = du.reshape(-1, H).T @ h_in.reshape(-1, h_in.shape[-1])
        net_db[i] = du.reshape(-1, H).sum(0)
        g = du @ p.net_w[i]
    dx = g                                                   # (B,P,in)
    draw = np.concatenate((dfreq * Z.dtype.type(15), dphase), 1)
    map_dW, map_db, dmap_in = mapping_backward(draw, p, t["mtape"])
    dZ = np.zeros_like(Z)
    if p.equivariance == "SO2":
        Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)
        D_xz = np.stack((D[:, :, 0], D[:, :, 2]), -1)
        d_ip = dx[:, :, 2:]
        dG = dmap_in[:, : N * N].reshape(B, N, N)
        dZ_xz = np.transpose(d_ip, (0, 2, 1)) @ D_xz + (dG + np.transpose(dG, (0, 2, 1))) @ Z_xz
        dZ[:, :, 0] = dZ_xz[:, :, 0]
        dZ[:, :, 2] = dZ_xz[:, :, 1]
        dZ[:, :, 1] = dmap_in[:, N * N:]
    else:
        dG = dmap_in.reshape(B, N, N)
        dZ = np.transpose(dx, (0, 2, 1)) @ D + (dG + np.transpose(dG, (0, 2, 1))) @ Z
    return dict(net_dW=net_dW, net_db=net_db, final_dW=final_dW, final_db=final_db, map_dW=map_dW, map_db=map_db,
                dZ=dZ, dfreq=dfreq, dphase=dphase)
