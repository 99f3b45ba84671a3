"""Environment-map Blinn-Phong shading: the downstream consumer of the decoder output on the reference's FIT_INVERSE path.

Mirrors ``EnvironmentMap`` and ``blinn_phong_shading_env_map`` of ``src/utils/pytorch3d_envmap_shader.py`` (:33-44,
:47-120; reached from ``RENI_module.get_render``, :386-396) from the point where the rasteriser's outputs are per-pixel
attributes -- the rasteriser itself (PyTorch3D) is out of scope.  Differentiable with respect to the environment map
(that is the gradient FIT_INVERSE sends back through ``model(Z, directions)``); the (B, H, W, J) and (B, H, W, J, 3)
tensors of the reference are never materialised.  CUDA only, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .functional import _call, _f32c, _require_cuda, _stream, _vp


class EnvironmentMap:
    """pytorch3d_envmap_shader.py:33-44: light directions and light colours = environment map x sine weight."""

    def __init__(self, environment_map: torch.Tensor = None, directions: torch.Tensor = None,
                 sineweight: torch.Tensor = None) -> None:
        self.directions = directions
        self.environment_map = environment_map * sineweight


class _ShadeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals, view, D, light, kd: float, ks: float, shininess: float):
        dev = _require_cuda(normals, view, D, light)
        n, v, lc = _f32c(normals), _f32c(view), _f32c(light)
        B, J = lc.shape[0], lc.shape[1]
        if D.shape[0] == 1 or D.stride(0) == 0:
            Dc, d_bs = _f32c(D[:1]), 0
        else:
            assert D.shape[0] == B, f"directions batch {D.shape[0]} != environment-map batch {B}"
            Dc, d_bs = _f32c(D), J * 3
        n_pix = n.shape[0]
        colors = torch.empty(B, n_pix, 3, device=dev, dtype=torch.float32)
        rc = _call(dev, _lib.load().reni_envmap_shade_forward, _vp(n), _vp(v), n_pix, _vp(Dc), d_bs, _vp(lc), B, J,
                   float(kd), float(ks), float(shininess), _vp(colors), _stream(dev))
        _lib.check(rc, "reni_envmap_shade_forward")
        ctx.save_for_backward(n, v, Dc)
        ctx.args = (d_bs, B, J, float(kd), float(ks), float(shininess))
        return colors

    @staticmethod
    def backward(ctx, grad_colors):
        n, v, Dc = ctx.saved_tensors
        d_bs, B, J, kd, ks, shininess = ctx.args
        dev = n.device
        g = _f32c(grad_colors)
        d_light = torch.empty(B, J, 3, device=dev, dtype=torch.float32)
        rc = _call(dev, _lib.load().reni_envmap_shade_backward, _vp(n), _vp(v), n.shape[0], _vp(Dc), d_bs, _vp(g), B, J,
                   kd, ks, shininess, _vp(d_light), _stream(dev))
        _lib.check(rc, "reni_envmap_shade_backward")
        return None, None, None, d_light, None, None, None


def blinn_phong_shading_env_map(pixel_normals: torch.Tensor, pixel_positions: torch.Tensor,
                                camera_position: torch.Tensor, envmap: EnvironmentMap, shininess, kd: float,
                                ks: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``pixel_normals`` / ``pixel_positions`` (H, W, 3): what ``interpolate_face_attributes`` gives the reference
    (:70-75, K = 1); ``camera_position`` (1, 3) or (3,); ``envmap.directions`` (B or 1, J, 3), ``envmap.environment_map``
    (B, J, 3).  Returns ``(colors (B, H, W, 3), pixel_normals (B, H, W, 3))`` like the reference (:119)."""
    H, W = pixel_normals.shape[:2]
    n = F.normalize(pixel_normals.reshape(-1, 3).float(), p=2, dim=-1, eps=1e-6)                          # :85
    v = F.normalize(camera_position.reshape(1, 3).float() - pixel_positions.reshape(-1, 3).float(), p=2, dim=-1,
                    eps=1e-6)                                                                             # :95-97
    s = float(shininess.reshape(-1)[0]) if isinstance(shininess, torch.Tensor) else float(shininess)
    light = envmap.environment_map
    colors = _ShadeFunction.apply(n, v, envmap.directions, light, float(kd), float(ks), s)
    B = light.shape[0]
    return colors.reshape(B, H, W, 3), n.reshape(1, H, W, 3).expand(B, H, W, 3)
