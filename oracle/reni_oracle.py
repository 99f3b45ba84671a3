"""CPU oracle for the RENI decoder hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (JADGardner/RENI,
``src/models/RENI.py``, ``src/utils/loss_functions.py``, ``src/utils/utils.py``).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it; the product (``reni_b200``) never
does and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports the reference modules from ``/root/reference`` in the build container, runs
them on seeded inputs and commits the inputs + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Two formulations are restated:

* the *direct* one, op for op as the reference computes it (build the
  ``(B, P, 2N+N^2+2)`` encoding, run the dense SIREN), used as the checker, and
* the *hoisted* one (per-map ``M_b``/``c_b``) that the CUDA path
  implements, so that the individual kernels (prologue, map-level backward) can be
  checked in isolation.  ``tests/test_oracle_golden.py`` proves direct == hoisted.

All functions take/return numpy arrays and work in the dtype of their inputs
(float64 for "truth", float32 to mimic the reference arithmetic).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# geometry helpers  (reference: src/utils/utils.py:46-91)
# --------------------------------------------------------------------------------------


def get_directions(sidelen: int, dtype=np.float32) -> np.ndarray:
    """Equirectangular unit directions, y-up.  (utils.py:46-65) -> (1, sidelen^2/2, 3)."""
    half = sidelen // 2
    # torch.linspace in the reference is float32; keep the same rounding when dtype is f32
    u = (np.linspace(1, sidelen, sidelen, dtype=dtype) - dtype(0.5)) / dtype(half)
    v = (np.linspace(1, half, half, dtype=dtype) - dtype(0.5)) / dtype(half)
    v_grid, u_grid = np.meshgrid(v, u, indexing="ij")
    theta = (dtype(np.pi) * (u_grid.reshape(-1) - dtype(1))).astype(dtype)
    phi = (dtype(np.pi) * v_grid.reshape(-1)).astype(dtype)
    d = np.stack(
        (np.sin(phi) * np.sin(theta), np.cos(phi), -np.sin(phi) * np.cos(theta)), -1
    ).astype(dtype)
    return d[None]


def get_sineweight(sidelen: int, dtype=np.float32) -> np.ndarray:
    """sin(polar angle) replicated over the 3 channels.  (utils.py:68-78) -> (1, P, 3)."""
    half = sidelen // 2
    v = (np.linspace(1, half, half, dtype=dtype) - dtype(0.5)) / dtype(half)
    phi = dtype(np.pi) * np.repeat(v, sidelen)
    sw = np.sin(phi).astype(dtype)
    return np.repeat(sw[:, None], 3, axis=1)[None]


def rectangle_mask(sidelen: int, row0: int, row1: int, col0: int, col1: int, dtype=np.float32):
    """Binary (1, P, 3) mask that is 1 inside rows [row0,row1) x cols [col0,col1).

    Stands in for ``get_mask`` (utils.py:81-91), which NEAREST-resizes a PNG; the
    shipped Mask-3 is a rectangle."""
    half = sidelen // 2
    m = np.zeros((half, sidelen), dtype=dtype)
    m[row0:row1, col0:col1] = 1
    return np.repeat(m.reshape(-1, 1), 3, axis=1)[None]


# --------------------------------------------------------------------------------------
# invariant encodings  (reference: src/models/RENI.py:23-60)
# --------------------------------------------------------------------------------------


def so2_invariant_representation(Z: np.ndarray, D: np.ndarray) -> np.ndarray:
    """RENI.py:31-53.  Z (B,N,3), D (B,P,3) -> (B,P,2N+N^2+2), columns [N, N^2, 1, N, 1]."""
    B, N, _ = Z.shape
    P = D.shape[1]
    Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)  # :37
    D_xz = np.stack((D[:, :, 0], D[:, :, 2]), -1)  # :38
    G = Z_xz @ np.transpose(Z_xz, (0, 2, 1))  # :40
    Z_xz_invar = np.broadcast_to(G.reshape(B, 1, N * N), (B, P, N * N))  # :42
    innerprod = D_xz @ np.transpose(Z_xz, (0, 2, 1))  # :44
    D_xz_norm = np.sqrt(D[:, :, 0] ** 2 + D[:, :, 2] ** 2)[:, :, None]  # :45
    Z_y = np.broadcast_to(Z[:, None, :, 1], (B, P, N))  # :47
    D_y = D[:, :, 1][:, :, None]  # :49
    return np.concatenate((innerprod, Z_xz_invar, D_xz_norm, Z_y, D_y), 2)  # :51


def so3_invariant_representation(Z: np.ndarray, D: np.ndarray) -> np.ndarray:
    """RENI.py:23-28 -> (B,P,N+N^2)."""
    B, N, _ = Z.shape
    P = D.shape[1]
    G = Z @ np.transpose(Z, (0, 2, 1))
    innerprod = D @ np.transpose(Z, (0, 2, 1))
    Z_invar = np.broadcast_to(G.reshape(B, 1, N * N), (B, P, N * N))
    return np.concatenate((innerprod, Z_invar), 2)


def no_invariance(Z: np.ndarray, D: np.ndarray) -> np.ndarray:
    """RENI.py:56-60 -> (B,P,N+3N)."""
    B, N, _ = Z.shape
    P = D.shape[1]
    innerprod = D @ np.transpose(Z, (0, 2, 1))
    Z_input = np.broadcast_to(Z.reshape(B, 1, N * 3), (B, P, N * 3))
    return np.concatenate((innerprod, Z_input), 2)


ENCODINGS = {
    "SO2": so2_invariant_representation,
    "SO3": so3_invariant_representation,
    "None": no_invariance,
}


def in_features(ndims: int, equivariance: str) -> int:
    """RENI.py:118-126."""
    if equivariance == "SO2":
        return 2 * ndims + ndims * ndims + 2
    if equivariance == "SO3":
        return ndims + ndims * ndims
    if equivariance == "None":
        return ndims * 3 + ndims
    raise ValueError(equivariance)


# --------------------------------------------------------------------------------------
# decoder  (reference: SineLayer RENI.py:63-87, net RENI.py:132-178)
# --------------------------------------------------------------------------------------


@dataclass
class DecoderParams:
    """Weights of ``RENIAutoDecoder.net``: ``weights[i]`` is (out,in) like nn.Linear."""

    weights: List[np.ndarray]
    biases: List[np.ndarray]
    first_omega_0: float = 30.0
    hidden_omega_0: float = 30.0
    last_layer_linear: bool = True
    output_activation: Optional[str] = "tanh"
    equivariance: str = "SO2"

    def astype(self, dtype) -> "DecoderParams":
        return DecoderParams(
            [w.astype(dtype) for w in self.weights],
            [b.astype(dtype) for b in self.biases],
            self.first_omega_0,
            self.hidden_omega_0,
            self.last_layer_linear,
            self.output_activation,
            self.equivariance,
        )

    @property
    def n_sine_layers(self) -> int:
        return len(self.weights) - (1 if self.last_layer_linear else 0)


def siren_init(
    rng: np.random.Generator,
    ndims: int,
    equivariance: str = "SO2",
    hidden_features: int = 256,
    hidden_layers: int = 5,
    out_features: int = 3,
    first_omega_0: float = 30.0,
    hidden_omega_0: float = 30.0,
    last_layer_linear: bool = True,
    output_activation: Optional[str] = "tanh",
    dtype=np.float32,
) -> DecoderParams:
    """Same distributions as the reference initialiser (RENI.py:76-84,157-160; biases keep
    nn.Linear's default U(+-1/sqrt(in))) -- not the same random stream as torch."""
    nin = in_features(ndims, equivariance)
    ws, bs = [], []
    fan = nin
    lim = 1.0 / nin
    ws.append(rng.uniform(-lim, lim, (hidden_features, nin)))
    bs.append(rng.uniform(-1 / np.sqrt(fan), 1 / np.sqrt(fan), hidden_features))
    for _ in range(hidden_layers):
        lim = np.sqrt(6 / hidden_features) / hidden_omega_0
        ws.append(rng.uniform(-lim, lim, (hidden_features, hidden_features)))
        bs.append(rng.uniform(-1 / np.sqrt(hidden_features), 1 / np.sqrt(hidden_features), hidden_features))
    lim = np.sqrt(6 / hidden_features) / hidden_omega_0
    ws.append(rng.uniform(-lim, lim, (out_features, hidden_features)))
    bs.append(rng.uniform(-1 / np.sqrt(hidden_features), 1 / np.sqrt(hidden_features), out_features))
    return DecoderParams(
        [w.astype(dtype) for w in ws],
        [b.astype(dtype) for b in bs],
        first_omega_0,
        hidden_omega_0,
        last_layer_linear,
        output_activation,
        equivariance,
    )


@dataclass
class ForwardTape:
    x: np.ndarray  # encoding (B,P,in)
    pre: List[np.ndarray] = field(default_factory=list)  # omega*(h W^T + b) for sine layers, y for linear
    act: List[np.ndarray] = field(default_factory=list)  # layer outputs
    out: Optional[np.ndarray] = None


def _omega(p: DecoderParams, i: int) -> float:
    return p.first_omega_0 if i == 0 else p.hidden_omega_0


def decoder_forward(Z: np.ndarray, D: np.ndarray, p: DecoderParams, tape: bool = False):
    """``RENIAutoDecoder.forward`` on latent codes (RENI.py:225-233): encoding then net."""
    dt = Z.dtype
    x = ENCODINGS[p.equivariance](Z, D).astype(dt)
    t = ForwardTape(x=x)
    h = x
    nl = len(p.weights)
    for i in range(nl):
        y = h @ p.weights[i].T + p.biases[i]
        is_sine = (i < nl - 1) or (not p.last_layer_linear)
        if is_sine:
            a = dt.type(_omega(p, i)) * y  # RENI.py:87
            h = np.sin(a)
            t.pre.append(a)
        else:
            h = y  # RENI.py:153-162
            t.pre.append(y)
        t.act.append(h)
    if p.output_activation == "tanh":  # RENI.py:175-176
        h = np.tanh(h)
    elif p.output_activation is not None:
        # "exp" raises in the reference (nn.Exp does not exist, RENI.py:173-174)
        raise AttributeError("module 'torch.nn' has no attribute 'Exp'")
    t.out = h
    return (h, t) if tape else h


def decoder_backward(Z: np.ndarray, D: np.ndarray, p: DecoderParams, t: ForwardTape, grad_out: np.ndarray):
    """Hand-derived reverse pass of ``decoder_forward`` (what autograd does for the
    reference).  Returns (dWs, dbs, dZ)."""
    B, N, _ = Z.shape
    dt = Z.dtype
    nl = len(p.weights)
    g = grad_out
    if p.output_activation == "tanh":
        g = g * (1 - t.out**2)
    dWs = [None] * nl
    dbs = [None] * nl
    for i in reversed(range(nl)):
        is_sine = (i < nl - 1) or (not p.last_layer_linear)
        if is_sine:
            g = g * np.cos(t.pre[i]) * dt.type(_omega(p, i))
        h_in = t.x if i == 0 else t.act[i - 1]
        g2 = g.reshape(-1, g.shape[-1])
        dWs[i] = g2.T @ h_in.reshape(-1, h_in.shape[-1])
        dbs[i] = g2.sum(0)
        g = g @ p.weights[i]
    dx = g  # (B,P,in_features)
    dZ = np.zeros_like(Z)
    if p.equivariance == "SO2":
        Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)
        D_xz = np.stack((D[:, :, 0], D[:, :, 2]), -1)
        d_ip = dx[:, :, :N]
        dG = dx[:, :, N : N + N * N].sum(1).reshape(B, N, N)
        dZy = dx[:, :, N + N * N + 1 : N + N * N + 1 + N].sum(1)
        dZ_xz = np.transpose(d_ip, (0, 2, 1)) @ D_xz + (dG + np.transpose(dG, (0, 2, 1))) @ Z_xz
        dZ[:, :, 0] = dZ_xz[:, :, 0]
        dZ[:, :, 2] = dZ_xz[:, :, 1]
        dZ[:, :, 1] = dZy
    elif p.equivariance == "SO3":
        d_ip = dx[:, :, :N]
        dG = dx[:, :, N:].sum(1).reshape(B, N, N)
        dZ = np.transpose(d_ip, (0, 2, 1)) @ D + (dG + np.transpose(dG, (0, 2, 1))) @ Z
    else:
        d_ip = dx[:, :, :N]
        dZ = np.transpose(d_ip, (0, 2, 1)) @ D + dx[:, :, N:].sum(1).reshape(B, N, 3)
    return dWs, dbs, dZ


# --------------------------------------------------------------------------------------
# losses  (reference: src/utils/loss_functions.py:6-71)
# --------------------------------------------------------------------------------------


def weighted_mse(o, t, sw):
    """loss_functions.py:6-13: sum_b mean_{p,c}((o-t)^2 * sw)."""
    B = o.shape[0]
    return (((o - t) ** 2) * sw).reshape(B, -1).mean(1).sum(0)


def weighted_cosine_similarity(o, t, sw, eps=1e-20):
    """loss_functions.py:25-32.  cosine_similarity over dim=1 (the PIXEL axis) -> (B,3),
    times sw[:,0] (the weight of pixel 0, (B,3)), mean over channels, sum_b (1 - .)."""
    num = (o * t).sum(1)
    # torch >= 1.12 clamps each norm separately: x.y / (max(|x|,eps) * max(|y|,eps))
    den = np.maximum(np.sqrt((o * o).sum(1)), eps) * np.maximum(np.sqrt((t * t).sum(1)), eps)
    cs = num / den
    return (1 - (cs * sw[:, 0]).mean(1)).sum(0)


def kld(mu, log_var, Z_dims=1):
    """loss_functions.py:16-22."""
    B = mu.shape[0]
    k = -0.5 * (1 + log_var - mu**2 - np.exp(log_var)).reshape(B, -1).sum(1)
    return (k / Z_dims).sum(0)


def reni_train_loss(o, t, sw):
    """RENITrainLoss (loss_functions.py:39-45)."""
    return weighted_mse(o, t, sw)


def reni_vad_train_loss(o, t, sw, mu, log_var, beta=1.0, Z_dims=None):
    """RENIVADTrainLoss (loss_functions.py:47-58)."""
    mse = weighted_mse(o, t, sw)
    k = beta * kld(mu, log_var, Z_dims)
    return mse + k, mse, k


def reni_test_loss(o, t, sw, Z, alpha=1.0, beta=1.0):
    """RENITestLoss (loss_functions.py:60-71) -> (loss, mse, prior, cosine)."""
    mse = weighted_mse(o, t, sw)
    prior = alpha * (Z**2).sum()
    cos = beta * weighted_cosine_similarity(o, t, sw)
    return mse + prior + cos, mse, prior, cos


def loss_grad_wrt_output(o, t, sw, beta=0.0, eps=1e-20):
    """d(mse + beta*cosine)/d o -- the g_o of SURVEY.md section 8a."""
    B, P, C = o.shape
    g = 2 * (o - t) * sw / (C * P)
    if beta != 0.0:
        no = np.maximum(np.sqrt((o * o).sum(1, keepdims=True)), eps)
        nt = np.maximum(np.sqrt((t * t).sum(1, keepdims=True)), eps)
        dot = (o * t).sum(1, keepdims=True)
        dcs = t / (no * nt) - dot * o / (no**3 * nt)
        g = g + beta * (-(sw[:, :1, :]) / C) * dcs
    return g


# --------------------------------------------------------------------------------------
# hoisted formulation -- what the CUDA path computes
# --------------------------------------------------------------------------------------


def split_w0_so2(W0: np.ndarray, N: int):
    """Column blocks of the first-layer weight in the order of RENI.py:51."""
    o = 0
    W_ip = W0[:, o : o + N]
    o += N
    W_G = W0[:, o : o + N * N]
    o += N * N
    w_dn = W0[:, o]
    o += 1
    W_zy = W0[:, o : o + N]
    o += N
    w_dy = W0[:, o]
    return W_ip, W_G, w_dn, W_zy, w_dy


def direction_features(D: np.ndarray, equivariance: str = "SO2") -> np.ndarray:
    """f = [d_x, d_z, |d_xz|, d_y] (SO2) or d (SO3/None), padded to 4."""
    if equivariance == "SO2":
        return np.stack(
            (D[..., 0], D[..., 2], np.sqrt(D[..., 0] ** 2 + D[..., 2] ** 2), D[..., 1]), -1
        )
    z = np.zeros_like(D[..., 0])
    return np.stack((D[..., 0], D[..., 1], D[..., 2], z), -1)


def hoist_layer0(Z: np.ndarray, W0: np.ndarray, b0: np.ndarray, equivariance: str = "SO2"):
    """Per-map M_b (B,4,H) and c_b (B,H) with a0 = omega * (f @ M_b + c_b)."""
    B, N, _ = Z.shape
    H = W0.shape[0]
    M = np.zeros((B, 4, H), dtype=Z.dtype)
    if equivariance == "SO2":
        W_ip, W_G, w_dn, W_zy, w_dy = split_w0_so2(W0, N)
        Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)
        G = Z_xz @ np.transpose(Z_xz, (0, 2, 1))
        c = G.reshape(B, N * N) @ W_G.T + Z[:, :, 1] @ W_zy.T + b0
        M[:, 0:2] = np.transpose(Z_xz, (0, 2, 1)) @ W_ip.T
        M[:, 2] = w_dn
        M[:, 3] = w_dy
    elif equivariance == "SO3":
        W_ip, W_G = W0[:, :N], W0[:, N:]
        G = Z @ np.transpose(Z, (0, 2, 1))
        c = G.reshape(B, N * N) @ W_G.T + b0
        M[:, 0:3] = np.transpose(Z, (0, 2, 1)) @ W_ip.T
    else:
        W_ip, W_Z = W0[:, :N], W0[:, N:]
        c = Z.reshape(B, N * 3) @ W_Z.T + b0
        M[:, 0:3] = np.transpose(Z, (0, 2, 1)) @ W_ip.T
    return M, c


def hoisted_forward(Z, D, p: DecoderParams):
    """Forward through the hoisted first layer; must equal ``decoder_forward``."""
    dt = Z.dtype
    M, c = hoist_layer0(Z, p.weights[0], p.biases[0], p.equivariance)
    f = direction_features(D, p.equivariance)
    h = np.sin(dt.type(p.first_omega_0) * (f @ M + c[:, None, :]))
    nl = len(p.weights)
    for i in range(1, nl):
        y = h @ p.weights[i].T + p.biases[i]
        is_sine = (i < nl - 1) or (not p.last_layer_linear)
        h = np.sin(dt.type(p.hidden_omega_0) * y) if is_sine else y
    if p.output_activation == "tanh":
        h = np.tanh(h)
    return h


def layer0_backward_so2(Z, W0, dM, dc):
    """Map-level backward: per-map dM (B,4,H), dc (B,H) ->
    (dW0 (H,in), db0 (H,), dZ (B,N,3))."""
    B, N, _ = Z.shape
    W_ip, W_G, w_dn, W_zy, w_dy = split_w0_so2(W0, N)
    Z_xz = np.stack((Z[:, :, 0], Z[:, :, 2]), -1)
    G = Z_xz @ np.transpose(Z_xz, (0, 2, 1))
    dW_ip = np.einsum("bch,bnc->hn", dM[:, 0:2], Z_xz)
    dW_G = np.einsum("bh,bk->hk", dc, G.reshape(B, N * N))
    dw_dn = dM[:, 2].sum(0)
    dW_zy = np.einsum("bh,bn->hn", dc, Z[:, :, 1])
    dw_dy = dM[:, 3].sum(0)
    db0 = dc.sum(0)
    dW0 = np.concatenate((dW_ip, dW_G, dw_dn[:, None], dW_zy, dw_dy[:, None]), 1)
    dG = (dc @ W_G).reshape(B, N, N)
    dZ_xz = (dG + np.transpose(dG, (0, 2, 1))) @ Z_xz + np.transpose(dM[:, 0:2] @ W_ip, (0, 2, 1))
    dZ_y = dc @ W_zy
    dZ = np.stack((dZ_xz[:, :, 0], dZ_y, dZ_xz[:, :, 1]), -1)
    return dW0, db0, dZ


# --------------------------------------------------------------------------------------
# one training / latent-fit step, as the reference's training_step does it
# (src/lightning/RENI_module.py:80-146; examples.ipynb cell 4)
# --------------------------------------------------------------------------------------


def step_fit_decoder(Z, D, target, sw, p: DecoderParams):
    """FIT_DECODER / AutoDecoder: RENITrainLoss, grads for every weight and for Z."""
    o, tape = decoder_forward(Z, D, p, tape=True)
    loss = reni_train_loss(o, target, sw)
    g = loss_grad_wrt_output(o, target, sw, beta=0.0)
    dWs, dbs, dZ = decoder_backward(Z, D, p, tape, g)
    return dict(out=o, loss=loss, dW=dWs, db=dbs, dZ=dZ)


def step_fit_latent(Z, D, target, sw, p: DecoderParams, alpha=1e-7, beta=1e-4):
    """FIT_LATENT: RENITestLoss (mse + alpha*sum Z^2 + beta*cosine), grads for Z only."""
    o, tape = decoder_forward(Z, D, p, tape=True)
    loss, mse, prior, cos = reni_test_loss(o, target, sw, Z, alpha, beta)
    g = loss_grad_wrt_output(o, target, sw, beta=beta)
    _, _, dZ = decoder_backward(Z, D, p, tape, g)
    dZ = dZ + 2 * alpha * Z
    return dict(out=o, loss=loss, mse_loss=mse, prior_loss=prior, cosine_loss=cos, dZ=dZ)


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def rel_max(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
