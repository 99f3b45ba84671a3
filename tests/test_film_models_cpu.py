"""Host-side behaviour of the FiLM drop-in modules (no GPU): state-dict layout and initialisation statistics against
fixtures taken from the reference's RENIAutoDecoderFiLM / RENIVariationalAutoDecoderFiLM (RENI.py:527-858), the per-map
operands the CUDA core consumes against the oracle, load_state_dict semantics, and the loud failure without CUDA."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O

import reni_film_oracle as FO  # noqa: E402
from reni_b200 import RENIAutoDecoderFiLM, RENIVariationalAutoDecoderFiLM
from test_film_oracle_golden import load_film_case

MOD = np.load(os.path.join(GOLDEN, "film_module.npz"))


def film_model_from_params(p, N, device="cpu", dataset_size=4, fixed=False, cls=RENIAutoDecoderFiLM):
    Lf, H = len(p.net_w), p.net_w[0].shape[0]
    m = cls(dataset_size, N, p.equivariance, H, Lf, p.map_w[0].shape[0], len(p.map_w) - 1, p.final_w.shape[0],
            p.output_activation, fixed)
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))  # noqa: E731
    with torch.no_grad():
        for i in range(Lf):
            m.net[i].layer.weight.copy_(f(p.net_w[i]))
            m.net[i].layer.bias.copy_(f(p.net_b[i]))
        m.final_layer.weight.copy_(f(p.final_w))
        m.final_layer.bias.copy_(f(p.final_b))
        for i in range(len(p.map_w)):
            m.mapping_network.network[2 * i].weight.copy_(f(p.map_w[i]))
            m.mapping_network.network[2 * i].bias.copy_(f(p.map_b[i]))
    return m.to(device)


def test_state_dict_keys_shapes_and_init_match_the_reference():
    torch.manual_seed(0)
    m = RENIAutoDecoderFiLM(7, 36, "SO2", 256, 5, 256, 3, 3, None, False)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in MOD["state_keys"]]
    for (k, v), shp, lo, hi, std in zip(sd.items(), MOD["state_shapes"], MOD["state_min"], MOD["state_max"], MOD["state_std"]):
        assert list(v.shape) == [int(x) for x in shp[: v.dim()]], k
        if k.endswith("weight") and (k.startswith("net.") or k.startswith("final_layer")):
            # uniform initialisers (RENI.py:463-478): same support
            bound = max(abs(lo), abs(hi))
            assert float(v.abs().max()) <= bound * 1.02 and float(v.abs().max()) > 0.9 * bound, k
        if k.startswith("mapping_network") and k.endswith("weight"):
            assert abs(float(v.std()) / std - 1) < 0.05, k  # kaiming normal, last layer x 0.25 (RENI.py:500-502)
    assert m.in_features == 38 and m.mn_in_features == 36 * 36 + 36


def test_vad_film_fixed_decoder_freezes_like_the_reference():
    mv = RENIVariationalAutoDecoderFiLM(7, 36, "SO2", 256, 5, 256, 3, 3, None, True)
    assert list(mv.state_dict().keys()) == [str(k) for k in MOD["vad_state_keys"]]
    assert [int(q.requires_grad) for q in mv.parameters()] == [int(x) for x in MOD["vad_fixed_requires_grad"]]
    assert float(mv.mu.abs().max()) == float(MOD["vad_fixed_mu_absmax"]) == 0.0
    Z, mu, lv = mv.sample_latent([1, 2])
    assert Z.shape == mu.shape == lv.shape == (2, 36, 3)


@pytest.mark.parametrize("name", ["film_so2_small", "film_so3_small_tanh", "film_so2_n9_h256"])
def test_map_level_operands_match_the_oracle(name):
    c = load_film_case(name)
    m = film_model_from_params(c["p"], c["N"]).double()
    mc, film = m.map_level(torch.from_numpy(c["Z"]).double())
    mc_o, film_o = FO.film_core_inputs(c["Z"].astype(np.float64), c["p"].astype(np.float64))
    assert O.rel_l2(mc.detach().numpy(), mc_o) < 1e-12
    assert O.rel_l2(film.detach().numpy(), film_o) < 1e-12


def test_load_state_dict_semantics():
    torch.manual_seed(3)
    src = RENIAutoDecoderFiLM(5, 4, "SO2", 256, 2, 16, 1, 3, None, False)
    ckpt = {"model." + k: v.clone() for k, v in src.state_dict().items()}
    ckpt["criterion.junk"] = torch.zeros(1)
    dst = RENIAutoDecoderFiLM(9, 4, "SO2", 256, 2, 16, 1, 3, None, True)   # fresh latents, frozen decoder
    dst.load_state_dict(ckpt)
    assert float(dst.Z.abs().max()) == 0.0 and dst.Z.shape[0] == 9
    for a, b in zip(src.core_parameters(), dst.core_parameters()):
        assert torch.equal(a, b) and not b.requires_grad
    assert torch.equal(src.mapping_network.network[0].weight, dst.mapping_network.network[0].weight)
    full = RENIAutoDecoderFiLM(5, 4, "SO2", 256, 2, 16, 1, 3, None, False)
    full.load_state_dict(ckpt, strict=False)
    assert torch.equal(full.Z, src.Z)


def test_no_cpu_fallback_and_reference_errors():
    m = RENIAutoDecoderFiLM(3, 4, "SO2", 256, 2, 16, 1, 3, None, False)
    D = torch.nn.functional.normalize(torch.randn(1, 8, 3), dim=-1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(0, D)
    with pytest.raises(NotImplementedError):
        m("bad", D)
    with pytest.raises(AssertionError):
        m([0, 1], D)
    none = RENIAutoDecoderFiLM(3, 4, "None", 256, 2, 16, 1, 3, None, False)
    with pytest.raises(RuntimeError, match="broken in the reference"):
        none(0, D)
    with pytest.raises(NotImplementedError):
        RENIAutoDecoderFiLM(3, 4, "SO2", 128, 2, 16, 1, 3, None, False)(0, D.cuda() if torch.cuda.is_available() else D)
