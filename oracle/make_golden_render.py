"""Golden vectors for the downstream consumer of the decoder output on the FIT_INVERSE path: the Blinn-Phong shading of a
surface lit by ALL texels of the environment map (reference: src/utils/pytorch3d_envmap_shader.py:47-120, called from
RENI_module.get_render, :386-396).  TEST INFRASTRUCTURE ONLY.

The reference function is run UNMODIFIED.  Its module imports PyTorch3D (absent here) for the rasteriser that produces the
per-pixel surface normals and positions; those are inputs of the shading, so PyTorch3D is stubbed: the fake
``interpolate_face_attributes`` hands back per-pixel attributes that the fixture supplies, everything from
``pixel_normals = F.normalize(...)`` on (lines 85-119) is the reference's own torch code.

    python oracle/make_golden_render.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("RENI_REFERENCE_PATH", "/root/reference")

# name: (seed, B maps, render H, render W, envmap sidelen, kd, shininess)
RENDER_CASES = {
    "render_small": (31, 2, 6, 5, 8, 0.5, 20.0),
    "render_64": (32, 2, 24, 24, 32, 0.5, 500.0),      # shininess 500 is the reference's material (build_renderer)
    "render_diffuse": (33, 1, 16, 16, 16, 1.0, 500.0),  # KD_VALUE = 1.0 (configs/default.py:83): no specular term
}


def render_inputs(seed, B, H, W, sidelen):
    """Deterministic fixture: unit-ish normals and surface points of a bumpy sphere cap seen from the +z camera, HDR-like
    positive light colours already multiplied by the sine weights (EnvironmentMap.__init__, :40-41)."""
    sys.path.insert(0, HERE)
    import reni_oracle as O

    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.linspace(-0.8, 0.8, H), np.linspace(-0.8, 0.8, W), indexing="ij")
    zs = np.sqrt(np.clip(1.0 - xs ** 2 - ys ** 2, 0.05, None))
    pos = np.stack((xs, ys, zs), -1).astype(np.float32)                                  # (H, W, 3)
    nrm = (pos + 0.15 * rng.standard_normal(pos.shape)).astype(np.float32) * 1.7          # NOT unit: the shader normalises
    cam = np.array([[0.0, 0.0, 2.0]], dtype=np.float32)                                   # look_at_view_transform(2.0, 0, 0)
    D = np.repeat(O.get_directions(sidelen), B, 0)                                        # (B, J, 3)
    sw = np.repeat(O.get_sineweight(sidelen), B, 0)
    env = np.exp(rng.uniform(-3.0, 2.0, (B, D.shape[1], 3))).astype(np.float32)           # unnormalised HDR radiance
    return pos, nrm, cam, D.astype(np.float32), sw.astype(np.float32), env


def _stub_pytorch3d():
    names = {
        "pytorch3d": [], "pytorch3d.structures": ["Meshes"], "pytorch3d.common": ["Device"],
        "pytorch3d.renderer": ["Materials", "TensorProperties", "look_at_view_transform", "FoVPerspectiveCameras",
                               "RasterizationSettings", "MeshRenderer", "MeshRasterizer", "TexturesVertex"],
        "pytorch3d.renderer.utils": ["TensorProperties"], "pytorch3d.ops": ["interpolate_face_attributes"],
        "pytorch3d.renderer.mesh": [], "pytorch3d.renderer.mesh.rasterizer": ["Fragments"], "pytorch3d.io": ["load_obj"],
        "pytorch3d.transforms": ["RotateAxisAngle"],
    }
    for mod, attrs in names.items():
        m = sys.modules.setdefault(mod, types.ModuleType(mod))
        for a in attrs:
            setattr(m, a, type(a, (), {}))
    # the rasteriser's interpolation: the fixture's "faces" are the pixels themselves, so the face attribute IS the
    # per-pixel attribute -> (N=1, H, W, K=1, 3)
    def interpolate_face_attributes(pix_to_face, bary_coords, face_attrs):
        H, W = pix_to_face
        return face_attrs.reshape(1, H, W, 1, 3)

    sys.modules["pytorch3d.ops"].interpolate_face_attributes = interpolate_face_attributes


def run_reference(pos, nrm, cam, D, sw, env, kd, shininess, dtype):
    import torch

    _stub_pytorch3d()
    sys.path.insert(0, REF)
    from src.utils import pytorch3d_envmap_shader as S

    tdt = torch.float32 if dtype == np.float32 else torch.float64
    H, W = pos.shape[:2]

    class Mesh:  # verts[faces] with faces = arange(HW)[:, None] -> (HW, 1, 3): one "vertex" per pixel
        def verts_packed(self): return torch.tensor(pos.reshape(-1, 3), dtype=tdt)
        def faces_packed(self): return torch.arange(H * W).reshape(-1, 1)
        def verts_normals_packed(self): return torch.tensor(nrm.reshape(-1, 3), dtype=tdt)

    class Frag:
        pix_to_face = (H, W)
        bary_coords = None

    class Cam:
        def get_camera_center(self): return torch.tensor(cam, dtype=tdt)

    class Mat:
        pass

    Mat.shininess = torch.tensor([shininess], dtype=tdt)
    e = torch.tensor(env, dtype=tdt, requires_grad=True)
    envmap = S.EnvironmentMap(environment_map=e, directions=torch.tensor(D, dtype=tdt), sineweight=torch.tensor(sw, dtype=tdt))
    colors, normals = S.blinn_phong_shading_env_map("cpu", Mesh(), Frag(), envmap, Cam(), Mat(), kd, 1.0 - kd)
    g = torch.tensor(np.random.default_rng(7).standard_normal(tuple(colors.shape)), dtype=tdt)
    (colors * g).sum().backward()
    return colors.detach().numpy(), normals.detach().numpy(), e.grad.numpy(), g.numpy()


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, (seed, B, H, W, sidelen, kd, shin) in RENDER_CASES.items():
        pos, nrm, cam, D, sw, env = render_inputs(seed, B, H, W, sidelen)
        store = {}
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            colors, normals, denv, g = run_reference(pos, nrm, cam, D, sw, env, kd, shin, dt)
            store[f"colors_{tag}"] = colors
            store[f"denv_{tag}"] = denv
            store[f"grad_out_{tag}"] = g
        store["normals_f32"] = normals.astype(np.float32)
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **store)
        print(name, store["colors_f32"].shape, float(np.abs(store["colors_f64"]).max()))


if __name__ == "__main__":
    main()
