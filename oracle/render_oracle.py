"""CPU oracle of the environment-map Blinn-Phong shading -- TEST INFRASTRUCTURE ONLY (see reni_oracle.py's header).

numpy restatement of ``blinn_phong_shading_env_map`` (src/utils/pytorch3d_envmap_shader.py:47-120) from the point where
the rasteriser's outputs are per-pixel attributes: surface normals and positions in, colours out, plus the gradient with
respect to the light colours (the environment map) that FIT_INVERSE back-propagates into the decoder
(RENI_module.py:107-112,386-396).  Pinned by tests/test_render_cpu.py against fixtures produced by running the
unmodified reference function (oracle/make_golden_render.py)."""
from __future__ import annotations

import numpy as np


def _normalize(x, eps=1e-6):
    """F.normalize(p=2, dim=-1, eps): x / max(||x||, eps)."""
    return x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), eps)


def shading_weights(normals, positions, camera, D, kd, ks, shininess):
    """(B, H, W, J) weight of every light on every pixel: kd * clamp(n.l) + c * ks * clamp(n.h)^s  (:85-118)."""
    n = _normalize(normals)                                                   # :85
    diffuse = np.clip(np.einsum("hwk,bjk->bhwj", n, D), 0.0, 1.0)             # :90-91
    v = _normalize(camera.reshape(1, 1, 3) - positions)                       # :95-97
    Hh = _normalize(v[None, :, :, None, :] + D[:, None, None, :, :])          # :105-106
    spec = np.clip(np.einsum("hwk,bhwjk->bhwj", n, Hh), 0.0, 1.0) ** shininess  # :108-110
    c = (shininess + 2.0) / (4.0 * (2.0 - np.exp(-shininess / 2.0)))          # :113-115
    return kd * diffuse + c * ks * spec, n


def blinn_phong_env_map(normals, positions, camera, D, light_colors, kd, ks, shininess):
    """colors (B, H, W, 3) = sum_j weights[b,h,w,j] * light_colors[b,j,:]  (:93,:112,:116), and the unit normals."""
    w, n = shading_weights(normals, positions, camera, D, kd, ks, shininess)
    return np.einsum("bjk,bhwj->bhwk", light_colors, w), n


def blinn_phong_env_map_backward(normals, positions, camera, D, grad_colors, kd, ks, shininess):
    """d sum(colors * grad_colors) / d light_colors -> (B, J, 3)."""
    w, _ = shading_weights(normals, positions, camera, D, kd, ks, shininess)
    return np.einsum("bhwj,bhwk->bjk", w, grad_colors)
