"""Direction / sine-weight / mask grids of the equirectangular parameterisation.

Mirrors ``get_directions`` / ``get_sineweight`` / ``get_mask`` of the reference's
``src/utils/utils.py:46-91`` (same formulas, same (1, P, 3) shapes, row-major over (v, u)).
These run once per resolution (RENI_module.py:43-44); they are plain torch and device-agnostic.
"""
from __future__ import annotations

import math

import torch


def _uv(sidelen: int):
    half = sidelen // 2
    u = (torch.linspace(1, sidelen, steps=sidelen) - 0.5) / half
    v = (torch.linspace(1, half, steps=half) - 0.5) / half
    v_grid, u_grid = torch.meshgrid(v, u, indexing="ij")
    return u_grid.reshape(-1), v_grid.reshape(-1)


def get_directions(sidelen: int) -> torch.Tensor:
    """Unit direction of every pixel of a (sidelen/2 x sidelen) panorama, y-up -> (1, P, 3).  (utils.py:46-65)"""
    u, v = _uv(sidelen)
    theta = math.pi * (u - 1)
    phi = math.pi * v
    return torch.stack(
        (torch.sin(phi) * torch.sin(theta), torch.cos(phi), -torch.sin(phi) * torch.cos(theta)), -1
    ).unsqueeze(0)


def get_sineweight(sidelen: int) -> torch.Tensor:
    """sin(polar angle) per pixel, replicated over RGB -> (1, P, 3).  (utils.py:68-78)"""
    _, v = _uv(sidelen)
    return torch.sin(math.pi * v).unsqueeze(1).repeat(1, 3).unsqueeze(0)


def get_mask(sidelen: int, path: str) -> torch.Tensor:
    """PNG mask -> NEAREST-resized (1, P, 3) tensor.  (utils.py:81-91; needs PIL + torchvision)"""
    from PIL import Image
    from torchvision import transforms

    mask = transforms.ToTensor()(Image.open(path))
    if mask.shape[0] == 1:
        mask = mask.repeat(3, 1, 1)
    mask = mask[:3]
    mask = transforms.Resize((sidelen // 2, sidelen), interpolation=transforms.InterpolationMode.NEAREST)(mask)
    return mask.permute((1, 2, 0)).reshape(-1, 3).unsqueeze(0)


def rectangle_mask(sidelen: int, row0: int, row1: int, col0: int, col1: int) -> torch.Tensor:
    """Binary (1, P, 3) mask, 1 inside rows [row0,row1) x cols [col0,col1): the geometry of the shipped
    in-painting masks (data/Masks/Mask-3.png is a rectangle), for synthetic benchmarks."""
    m = torch.zeros(sidelen // 2, sidelen)
    m[row0:row1, col0:col1] = 1
    return m.reshape(-1, 1).repeat(1, 3).unsqueeze(0)


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """Binary (1, P, 3) / (P, 3) / (P,) mask -> int32 words with one bit per pixel (bit p & 31 of word p >> 5), the form
    the kernels take with RENI_FLAG_GRID_SINEWEIGHT (the mask of RENI_module.py:92-94 applied inside the kernels).
    Raises if the mask is not binary or differs between the channels of a pixel (get_mask's PNGs do neither)."""
    m = mask.reshape(-1, mask.shape[-1]) if mask.dim() > 1 else mask.reshape(-1, 1)
    if not bool(((m == 0) | (m == 1)).all()) or not bool((m == m[:, :1]).all()):
        raise ValueError("in-kernel masks must be binary and equal on the three channels of a pixel")
    bits = (m[:, 0] != 0).to(torch.int64)
    P = bits.numel()
    pad = (-P) % 32
    if pad:
        bits = torch.cat([bits, bits.new_zeros(pad)])
    weights = (1 << torch.arange(32, dtype=torch.int64, device=bits.device))
    words = (bits.reshape(-1, 32) * weights).sum(dim=1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)  # two's complement into int32
    return words.to(torch.int32).contiguous()
