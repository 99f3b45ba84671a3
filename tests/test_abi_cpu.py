"""The C-ABI library loads on a CPU-only box and exports every symbol include/reni_b200.h declares;
host-side entry points that need no GPU behave (sizes, error codes)."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT

import __graft_entry__ as entry
from reni_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "reni_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(reni_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/reni_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in reni_b200/_lib.py"
    assert _lib.missing_symbols() == []


def test_version_and_strerror(lib):
    assert lib.reni_abi_version() == 2
    assert lib.reni_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.reni_strerror(code)) > 4


def cfg(**kw):
    d = dict(ndims=36, equivariance=1, hidden_features=256, hidden_layers=5, out_features=3, last_layer_linear=1,
             output_activation=1, first_omega_0=30.0, hidden_omega_0=30.0)
    d.update(kw)
    return _lib.RENIConfig(**d)


def test_in_features_matches_reference_formula(lib):
    # src/models/RENI.py:118-126
    for N in (9, 36, 49, 100):
        assert lib.reni_in_features(C.byref(cfg(ndims=N, equivariance=1))) == 2 * N + N * N + 2
        assert lib.reni_in_features(C.byref(cfg(ndims=N, equivariance=2))) == N + N * N
        assert lib.reni_in_features(C.byref(cfg(ndims=N, equivariance=0))) == 4 * N


def test_workspace_sizes(lib):
    c = cfg()
    inf = lib.reni_workspace_bytes(C.byref(c), 32, 8192, 0)
    lat = lib.reni_workspace_bytes(C.byref(c), 32, 8192, 1)
    full = lib.reni_workspace_bytes(C.byref(c), 32, 8192, 3)
    assert 0 < inf < lat < full
    ntiles = 32 * 64
    # latent-only: phase stash (6 images of 64 KB at 16 bits per phase, 48 KB at 12) per tile; full: phase + delta
    # (6 images of 64 KB) + g_y
    assert lat - inf >= ntiles * 6 * 49152
    assert full - inf >= ntiles * 6 * (49152 + 65536)
    assert full < 2.0e9
    # ragged P rounds up to whole 128-direction tiles
    assert lib.reni_workspace_bytes(C.byref(c), 1, 129, 1) > lib.reni_workspace_bytes(C.byref(c), 1, 128, 1)


def test_bad_configs_are_rejected(lib):
    assert lib.reni_workspace_bytes(C.byref(cfg(hidden_features=128)), 1, 128, 0) == -1
    assert lib.reni_workspace_bytes(C.byref(cfg(hidden_layers=0)), 1, 128, 0) == -1
    assert lib.reni_workspace_bytes(C.byref(cfg(hidden_layers=7)), 1, 128, 0) == -1
    assert lib.reni_workspace_bytes(C.byref(cfg(equivariance=3)), 1, 128, 0) == -1
    assert lib.reni_workspace_bytes(C.byref(cfg(out_features=4)), 1, 128, 0) == -1
    assert lib.reni_workspace_bytes(C.byref(cfg()), 0, 128, 0) == -2
    assert lib.reni_workspace_bytes(C.byref(cfg()), 1, 0, 0) == -2
    with pytest.raises(_lib.RENILibraryError):
        _lib.check(-3, "x")


def test_null_arguments_return_error_codes_without_touching_the_gpu(lib):
    c = cfg()
    assert lib.reni_forward(C.byref(c), None, None, 0, None, None, 1, 128, None, None, None, 0, None, 0, 0, None) == -2
    assert lib.reni_prepare_weights(C.byref(c), None, None, None, 0, None) == -2
    assert lib.reni_selftest_umma(None, 0, None, 0, 0, 0, 0, 0, 0, 0, 0, 0, 16, 1, None, None) == -2


def test_film_per_map_entry_points_host_side(lib):
    """Sizes and argument checks of the FiLM additions that need no GPU: per-map weight images
    (RENI_FLAG_FILM_PERMAP / reni_film_prepare_maps), the training per-map stage's activation buffer, overlap hook."""
    c = cfg(hidden_layers=4, first_omega_0=1.0, hidden_omega_0=1.0)
    FILM, SAVE, PERMAP = _lib.FLAG_FILM, _lib.FLAG_SAVE_FOR_BACKWARD, _lib.FLAG_FILM_PERMAP
    B, P, L = 32, 8192, 4
    plain = lib.reni_workspace_bytes(C.byref(c), B, P, FILM | SAVE)
    permap = lib.reni_workspace_bytes(C.byref(c), B, P, FILM | SAVE | PERMAP)
    # three fp16 weight images (forward hi + residual, backward layout) + one bias block pair per map and layer
    assert permap - plain >= B * L * (3 * 131072 + 2 * 4096)
    assert permap - plain < B * L * (3 * 131072 + 2 * 4096) + 8192
    # the per-map flag means nothing without the FiLM flag
    assert lib.reni_workspace_bytes(C.byref(c), B, P, SAVE | PERMAP) == lib.reni_workspace_bytes(C.byref(c), B, P, SAVE)
    # P must be a multiple of 512 (the four tiles of a CTA pair's unit share one map); NULLs are refused
    dummy = (C.c_void_p * 6)()
    one = C.c_void_p(1024)
    assert lib.reni_film_prepare_maps(C.byref(c), one, dummy, dummy, 2, 640, one, 1 << 30, FILM, None) == -2
    assert lib.reni_film_prepare_maps(C.byref(c), None, dummy, dummy, 2, 512, one, 1 << 30, FILM, None) == -2
    dims = (C.c_int32 * 5)(36 * 36 + 36, 256, 256, 256, 2 * 5 * 256)
    acts = lib.reni_film_map_acts_bytes(dims, 4, 32)
    assert acts == 32 * 4 * sum(dims) + 256  # every activation kept (multiples of 256 bytes here) + the barrier words
    assert lib.reni_film_map_acts_bytes(dims, 0, 32) == -2
    assert lib.reni_film_map_acts_bytes(dims, 4, 0) == -2
    assert lib.reni_film_map_forward_train(C.byref(c), None, None, None, None, None, dims, 4, 32, None, None, None, 0,
                                           None) == -2
    assert lib.reni_film_map_backward(C.byref(c), None, None, None, None, dims, 4, 32, None, None, None, None, None,
                                      None, None, None, None, 0, None) == -2
    assert lib.reni_debug_set_overlap(-2, 0) == -2
    assert lib.reni_debug_set_overlap(0, 0) == 0
