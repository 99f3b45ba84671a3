"""The shading oracle against fixtures produced by the unmodified reference function
(src/utils/pytorch3d_envmap_shader.py:47-120 run by oracle/make_golden_render.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, ROOT  # noqa: F401  (path setup)
import render_oracle as RO
from make_golden_render import RENDER_CASES, render_inputs


@pytest.mark.parametrize("name", list(RENDER_CASES))
def test_shading_oracle_matches_reference_golden(name):
    seed, B, H, W, sidelen, kd, shin = RENDER_CASES[name]
    pos, nrm, cam, D, sw, env = render_inputs(seed, B, H, W, sidelen)
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    f64 = lambda a: a.astype(np.float64)  # noqa: E731
    light = f64(env) * f64(sw)  # EnvironmentMap.__init__ (:40-41)
    colors, n = RO.blinn_phong_env_map(f64(nrm), f64(pos), f64(cam), f64(D), light, kd, 1.0 - kd, shin)
    np.testing.assert_allclose(colors, g["colors_f64"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.broadcast_to(n[None], g["normals_f32"].shape), g["normals_f32"], atol=1e-6)
    d_light = RO.blinn_phong_env_map_backward(f64(nrm), f64(pos), f64(cam), f64(D), g["grad_out_f64"], kd, 1.0 - kd, shin)
    # the reference differentiates w.r.t. the environment map BEFORE the sine weights: d env = d light * sw
    np.testing.assert_allclose(d_light * f64(sw), g["denv_f64"], rtol=1e-10, atol=1e-12)
    # fp32 reference vs fp64 truth: the level any fp32 implementation can be held to (x^500 turns an fp32 rounding of
    # x into 500 * 6e-8 = 3e-5 of the term)
    assert np.abs(g["colors_f32"] - g["colors_f64"]).max() <= 2e-4 * np.abs(g["colors_f64"]).max()
