"""Host-side mirror of the reference module API (no GPU): constructor, parameter names/shapes,
init ranges, fixed-decoder semantics, state-dict loading, dispatch validation, geometry, losses."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, load_case

import reni_b200
from reni_b200 import (KLD, RENIAutoDecoder, RENITestLoss, RENITrainLoss, RENIVADTrainLoss, RENIVariationalAutoDecoder,
                       get_directions, get_model, get_sineweight, rectangle_mask, shard_range)

MOD = np.load(os.path.join(GOLDEN, "module.npz"))


def make(N=36, fixed=False, cls=RENIAutoDecoder, ds=7, **kw):
    a = dict(equivariance="SO2", hidden_features=256, hidden_layers=5, out_features=3, last_layer_linear=True,
             output_activation="tanh", first_omega_0=30.0, hidden_omega_0=30.0)
    a.update(kw)
    return cls(ds, N, a["equivariance"], a["hidden_features"], a["hidden_layers"], a["out_features"],
               a["last_layer_linear"], a["output_activation"], a["first_omega_0"], a["hidden_omega_0"], fixed)


def test_state_dict_matches_reference_keys_shapes_and_init_ranges():
    torch.manual_seed(0)
    m = make()
    sd = m.state_dict()
    assert list(sd.keys()) == list(MOD["state_keys"])
    for k, shp, lo, hi in zip(MOD["state_keys"], MOD["state_shapes"], MOD["state_min"], MOD["state_max"]):
        v = sd[str(k)]
        assert list(v.shape) == [int(s) for s in shp[: v.dim()]]
        if str(k) != "Z":  # uniform init: same support as the reference's constructor (RENI.py:76-84,157-160)
            bound = max(abs(lo), abs(hi))
            assert 0.9 * bound < float(v.abs().max()) <= bound * 1.001 + 1e-9
    assert sum(p.numel() for p in m.net.parameters()) == int(MOD["n_net_params"]) == 680707
    for n_ in (9, 49, 100):
        assert sum(p.numel() for p in make(N=n_, ds=1).net.parameters()) == int(MOD[f"n_net_params_{n_}"])
    assert m.in_features == 2 * 36 + 36 * 36 + 2


def test_fixed_decoder_freezes_net_and_zero_inits_latents():
    m = make(fixed=True)
    assert float(m.Z.abs().max()) == float(MOD["fixed_Z_absmax"]) == 0.0
    assert [int(p.requires_grad) for p in m.net.parameters()] == list(MOD["fixed_requires_grad"])
    assert m.Z.requires_grad
    v = make(fixed=True, cls=RENIVariationalAutoDecoder)
    assert float(v.mu.abs().max()) == 0.0 and not v.log_var.requires_grad and v.mu.requires_grad
    v2 = make(fixed=False, cls=RENIVariationalAutoDecoder)
    assert v2.log_var.requires_grad and abs(float(v2.log_var.mean()) + 5) < 0.2  # N(-5, 1), RENI.py:338-340


def test_load_state_dict_strips_lightning_prefix_and_respects_fixed_decoder():
    src = make(ds=3)
    ckpt = {"model." + k: v.clone() for k, v in src.state_dict().items()}
    ckpt["not_model.junk"] = torch.zeros(1)
    dst = make(ds=3)
    dst.load_state_dict(ckpt)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v)
    fx = make(ds=5, fixed=True)  # different dataset size: only net.* is loaded (RENI.py:196-201)
    fx.load_state_dict(ckpt)
    assert torch.equal(fx.net[3].linear.weight, src.net[3].linear.weight)
    assert float(fx.Z.abs().max()) == 0.0


def test_forward_validation_mirrors_reference():
    m = make(N=4, ds=6)
    D = torch.zeros(2, 16, 3)
    with pytest.raises(AssertionError):
        m([1, 2, 3], D)          # len(idx) != directions.shape[0]  (RENI.py:220)
    with pytest.raises(AssertionError):
        m(3, D)                  # RENI.py:213
    with pytest.raises(NotImplementedError):
        m("3", D)                # RENI.py:205-209
    with pytest.raises(AttributeError):
        make(output_activation="exp")   # nn.Exp does not exist (RENI.py:173-174)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 4, 3), D)
    with pytest.raises(NotImplementedError):
        make(N=4, hidden_features=512)(torch.zeros(2, 4, 3), D)  # wider than the kernels' 256 features
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        make(N=4, hidden_features=64)(torch.zeros(2, 4, 3), D)   # narrower decoders run zero-padded (on the GPU)


def test_vad_sample_latent_is_reparameterised():
    torch.manual_seed(3)
    v = make(N=4, ds=5, cls=RENIVariationalAutoDecoder)
    s, mu, lv = v.sample_latent([1, 3])
    assert s.shape == (2, 4, 3) and s.requires_grad
    assert torch.allclose(mu, v.mu[[1, 3]])
    assert float((s - mu).abs().max()) < 6 * float(torch.exp(0.5 * lv).max())


def test_get_model_factory():
    def cfg(model_type, cond="Cond-by-Concat"):
        r = types.SimpleNamespace(CONDITIONING=cond, LATENT_DIMENSION=9, EQUIVARIANCE="SO2", HIDDEN_FEATURES=256,
                                  HIDDEN_LAYERS=5, OUT_FEATURES=3, LAST_LAYER_LINEAR=True, OUTPUT_ACTIVATION="tanh",
                                  FIRST_OMEGA_0=30.0, HIDDEN_OMEGA_0=30.0, MAPPING_LAYERS=3, MAPPING_FEATURES=256,
                                  MODEL_TYPE=model_type)
        return types.SimpleNamespace(RENI=r)
    m = get_model(cfg("AutoDecoder"), 11, "FIT_DECODER")
    assert isinstance(m, RENIAutoDecoder) and not m.fixed_decoder and m.Z.shape == (11, 9, 3)
    m = get_model(cfg("VariationalAutoDecoder"), 11, "FIT_LATENT")
    assert isinstance(m, RENIVariationalAutoDecoder) and m.fixed_decoder   # RENI.py:874
    from reni_b200 import RENIAutoDecoderFiLM, RENIVariationalAutoDecoderFiLM
    m = get_model(cfg("AutoDecoder", "FiLM"), 11, "FIT_DECODER")       # RENI.py:905-917
    assert isinstance(m, RENIAutoDecoderFiLM) and not m.fixed_decoder and m.Z.shape == (11, 9, 3)
    assert m.siren_hidden_layers == 5 and m.mapping_network_layers == 3 and m.output_activation == "tanh"
    m = get_model(cfg("VariationalAutoDecoder", "FiLM"), 11, "FIT_INVERSE")
    assert isinstance(m, RENIVariationalAutoDecoderFiLM) and m.fixed_decoder and float(m.mu.abs().max()) == 0.0


def test_geometry_matches_reference_fixtures():
    g = np.load(os.path.join(GOLDEN, "geometry.npz"))
    for W in (16, 32):
        assert get_directions(W).shape == (1, W * W // 2, 3)
        np.testing.assert_allclose(get_directions(W).numpy(), g[f"dir_{W}"], atol=2e-7)
        np.testing.assert_allclose(get_sineweight(W).numpy(), g[f"sw_{W}"], atol=2e-7)
    d = get_directions(128).double()
    np.testing.assert_allclose([float(d.sum()), float(d.abs().sum()), float((d ** 2).sum())], g["dir_128_checksum"],
                               rtol=1e-6, atol=1e-3)
    m = rectangle_mask(128, 10, 46, 40, 82)
    assert m.shape == (1, 8192, 3) and abs(float(m.mean()) - 36 * 42 / 8192) < 1e-6


@pytest.mark.parametrize("name", ["so2_small", "so2_n36_h256_masked"])
def test_torch_losses_match_reference_values(name):
    c = load_case(name)
    g = c["g"]
    o = torch.from_numpy(g["out_f64"])
    t = torch.from_numpy(c["target"]).double()
    sw = torch.from_numpy(c["sw"]).double()
    Z = torch.from_numpy(c["Z"]).double()
    assert abs(float(RENITrainLoss()(o, t, sw)) - float(g["train_loss_f64"])) < 1e-12
    vals = [float(v) for v in RENITestLoss(alpha=c["alpha"], beta=c["beta"])(o, t, sw, Z)]
    np.testing.assert_allclose(vals, g["test_loss_f64"], rtol=1e-10)


def test_kld_and_vad_loss_match_reference():
    mu, lv = torch.from_numpy(MOD["kld_mu"]), torch.from_numpy(MOD["kld_lv"])
    assert abs(float(KLD(mu, lv, Z_dims=12)) - float(MOD["kld_val"])) < 1e-5 * abs(float(MOD["kld_val"]))
    o, t, sw = torch.rand(3, 8, 3), torch.rand(3, 8, 3), torch.rand(3, 8, 3)
    loss, mse, kl = RENIVADTrainLoss(beta=1e-4, Z_dims=12)(o, t, sw, mu, lv)
    assert abs(float(loss) - float(mse) - float(kl)) < 1e-7


def test_shard_range_partitions_maps():
    for n, w in ((32, 8), (33, 8), (5, 8), (256, 4), (4096, 3)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_multires_curriculum_hook_doubles_resolution_like_the_callback():
    """callbacks.py:11-25 with the default schedule (configs/default.py: 16x32 -> 64x128 at epochs 25, 80[, 150])."""
    from reni_b200 import RENITrainer

    m = make(N=4, ds=3, fixed=True)
    mask_fn = lambda w: rectangle_mask(w, w // 8, w // 4, w // 4, w // 2)  # noqa: E731
    tr = RENITrainer(m, "FIT_LATENT", 32, mask=mask_fn(32))
    curriculum = [25, 80]
    changed = [e for e in range(100) if tr.on_train_epoch_end(e, curriculum, mask_fn)]
    assert changed == [24, 79]                       # current_epoch + 1 in curriculum
    assert tr.sidelen == 128 and tr.directions.shape == (1, 8192, 3) and tr.sineweight.shape == (1, 8192, 3)
    assert tr.mask.shape == (1, 8192, 3)
    with pytest.raises(ValueError):
        tr.on_train_epoch_end(24, curriculum)        # masked task without a way to rebuild the mask


def test_get_mask_nearest_resize(tmp_path):
    """get_mask (utils.py:81-91): PNG -> ToTensor -> NEAREST resize -> (1, P, 3).  Checked against direct index sampling
    of a synthetic 512x256 mask (the geometry of data/Masks/*.png) and, where the reference is mounted, against the
    reference's own get_mask on the same file."""
    import sys
    import types

    from PIL import Image

    from reni_b200 import get_mask

    rng = np.random.default_rng(0)
    src = np.zeros((256, 512), dtype=np.uint8)
    src[40:186, 162:329] = 255                       # Mask-3-like rectangle
    src[rng.integers(0, 256, 300), rng.integers(0, 512, 300)] = 255   # plus isolated pixels
    for mode, name in (("L", "m1.png"), ("RGB", "m3.png")):
        arr = src if mode == "L" else np.repeat(src[:, :, None], 3, 2)
        path = str(tmp_path / name)
        Image.fromarray(arr, mode=mode).save(path)
        for W in (32, 128):
            got = get_mask(W, path)
            assert got.shape == (1, W * W // 2, 3)
            rows = (np.arange(W // 2) * (256 / (W // 2))).astype(np.int64)
            cols = (np.arange(W) * (512 / W)).astype(np.int64)
            want = (src[np.ix_(rows, cols)] / 255.0).astype(np.float32).reshape(-1, 1).repeat(3, 1)[None]
            np.testing.assert_array_equal(got.numpy(), want)
            if os.path.isdir("/root/reference/src"):
                sys.modules.setdefault("gdown", types.ModuleType("gdown"))
                if "/root/reference" not in sys.path:
                    sys.path.insert(0, "/root/reference")
                from src.utils import utils as ref_utils
                np.testing.assert_array_equal(got.numpy(), ref_utils.get_mask(W, path).numpy())
