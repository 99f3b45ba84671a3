"""Drop-in RENI decoder / auto-decoder modules backed by the sm_100a kernels.

Same class names, constructor arguments, attributes, parameter names and ``forward`` dispatch as
the reference's ``src/models/RENI.py`` (RENIAutoDecoder :90-233, RENIVariationalAutoDecoder
:236-399, SineLayer :63-87, get_model :861-933), so that the reference's training /
latent-optimisation loops (src/lightning/RENI_module.py:80-146, examples.ipynb cell 4) and its
checkpoints (state-dict keys ``Z`` / ``mu`` / ``log_var`` / ``net.{i}.linear.{weight,bias}`` /
``net.{L+1}.{weight,bias}``) work unchanged.  ``self.net`` only HOLDS the parameters: the
arithmetic of ``self.net(self.InvariantRepresentation(Z, D))`` runs in libreni_b200.so
(``reni_b200.functional.decode``).  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import functional as F_
from .functional import DecoderSpec, Workspace


class SineLayer(nn.Module):
    """Parameter holder with the reference's SIREN initialisation (RENI.py:63-84).

    first layer: W ~ U(+-1/in); others: W ~ U(+-sqrt(6/in)/omega_0); bias keeps nn.Linear's default.
    ``forward`` exists for API compatibility but the decoder never calls it -- the fused kernels
    read ``linear.weight`` / ``linear.bias`` directly."""

    def __init__(self, in_features, out_features, bias=True, is_first=False, omega_0=30):
        super().__init__()
        self.omega_0 = omega_0
        self.is_first = is_first
        self.in_features = in_features
        self.linear = nn.Linear(in_features, out_features, bias=bias)
        self.init_weights()

    def init_weights(self):
        with torch.no_grad():
            if self.is_first:
                bound = 1 / self.in_features
            else:
                bound = math.sqrt(6 / self.in_features) / self.omega_0
            self.linear.weight.uniform_(-bound, bound)

    def forward(self, input):  # pragma: no cover - not on the product path
        raise RuntimeError("SineLayer is evaluated inside the fused reni_b200 decoder kernel; call the decoder module")


def in_features_for(ndims: int, equivariance: str) -> int:
    """RENI.py:118-126."""
    if equivariance == "None":
        return ndims * 3 + ndims
    if equivariance == "SO2":
        return 2 * ndims + ndims * ndims + 2
    if equivariance == "SO3":
        return ndims + ndims * ndims
    raise ValueError(f"unknown equivariance {equivariance!r}")


class _DecoderBase(nn.Module):
    """Shared by the auto-decoder and the variational auto-decoder: hyper-parameters, ``net``, dispatch."""

    def __init__(self, dataset_size, ndims, equivariance, hidden_features, hidden_layers, out_features,
                 last_layer_linear, output_activation, first_omega_0, hidden_omega_0, fixed_decoder):
        super().__init__()
        self.dataset_size = dataset_size
        self.ndims = ndims
        self.equivariance = equivariance
        self.hidden_features = hidden_features
        self.hidden_layers = hidden_layers
        self.out_features = out_features
        self.last_layer_linear = last_layer_linear
        self.output_activation = output_activation
        self.first_omega_0 = first_omega_0
        self.hidden_omega_0 = hidden_omega_0
        self.fixed_decoder = fixed_decoder
        self.in_features = in_features_for(ndims, equivariance)

        self.init_latent_codes(dataset_size, ndims, fixed_decoder)

        net: List[nn.Module] = [SineLayer(self.in_features, hidden_features, is_first=True, omega_0=first_omega_0)]
        for _ in range(hidden_layers):
            net.append(SineLayer(hidden_features, hidden_features, is_first=False, omega_0=hidden_omega_0))
        if last_layer_linear:
            final_linear = nn.Linear(hidden_features, out_features)
            with torch.no_grad():
                bound = math.sqrt(6 / hidden_features) / hidden_omega_0
                final_linear.weight.uniform_(-bound, bound)
            net.append(final_linear)
        else:
            net.append(SineLayer(hidden_features, out_features, is_first=False, omega_0=hidden_omega_0))
        if output_activation == "exp":
            net.append(nn.Exp())  # AttributeError, exactly like the reference (RENI.py:173-174)
        elif output_activation == "tanh":
            net.append(nn.Tanh())
        self.net = nn.Sequential(*net)

        if fixed_decoder:
            for param in self.net.parameters():
                param.requires_grad = False

        self._ws = Workspace()  # inference workspace (differentiated forwards own theirs)

    # ---- plumbing -------------------------------------------------------------------------
    @property
    def spec(self) -> DecoderSpec:
        return DecoderSpec(self.ndims, self.equivariance, self.hidden_features, self.hidden_layers,
                           self.out_features, bool(self.last_layer_linear), self.output_activation,
                           float(self.first_omega_0), float(self.hidden_omega_0))

    def decoder_parameters(self) -> List[torch.Tensor]:
        """[W0, b0, ..., W_out, b_out] in state-dict order."""
        ps: List[torch.Tensor] = []
        for i in range(self.hidden_layers + 2):
            layer = self.net[i]
            lin = layer.linear if isinstance(layer, SineLayer) else layer
            ps += [lin.weight, lin.bias]
        return ps

    def decoder_weights(self) -> List[torch.Tensor]:
        return self.decoder_parameters()[0::2]

    def decoder_biases(self) -> List[torch.Tensor]:
        return self.decoder_parameters()[1::2]

    def decode(self, Z: torch.Tensor, directions: torch.Tensor) -> torch.Tensor:
        """net(InvariantRepresentation(Z, directions)) -> (B, P, out_features), fused on the GPU."""
        return F_.decode(self.spec, self._ws, Z, directions, self.decoder_parameters())

    def load_state_dict(self, state_dict, strict: bool = True):
        """Reference semantics (RENI.py:190-203): keep only keys prefixed ``model.`` (Lightning checkpoints), strip
        the prefix, and with ``fixed_decoder`` load the decoder weights only, leaving the fresh latents."""
        new_state_dict = {k[6:]: v for k, v in state_dict.items() if k.startswith("model.")}
        if self.fixed_decoder:
            net_state_dict = {k[4:]: v for k, v in new_state_dict.items() if k.startswith("net.")}
            return self.net.load_state_dict(net_state_dict, strict=strict)
        return super().load_state_dict(new_state_dict, strict=strict)

    # ---- forward dispatch (RENI.py:205-233 / :362-399) ------------------------------------------
    def _latents_for(self, idx):
        raise NotImplementedError

    def forward(self, x, directions):
        if not isinstance(x, (int, list, torch.Tensor)):
            raise NotImplementedError(
                "x must be either an int (idx), torch.Tensor (idxs or latent codes) or a list of ints (idxs)")
        if isinstance(x, int):
            assert len([x]) == directions.shape[0]
            Z = self._latents_for([x])
        elif isinstance(x, list):
            assert len(x) == directions.shape[0]
            Z = self._latents_for(x)
        elif len(x.shape) == 1:
            Z = self._latents_for(x)
        else:
            Z = x
        return self.decode(Z, directions)


class RENIAutoDecoder(_DecoderBase):
    """Reference: src/models/RENI.py:90-233."""

    def init_latent_codes(self, dataset_size, ndims, fixed_decoder=False):
        if fixed_decoder:
            self.Z = nn.Parameter(torch.zeros(dataset_size, ndims, 3))
        else:
            self.Z = nn.Parameter(torch.randn((dataset_size, ndims, 3)))

    def _latents_for(self, idx):
        return self.Z[idx, :, :]


class RENIVariationalAutoDecoder(_DecoderBase):
    """Reference: src/models/RENI.py:236-399.  Sampling stays in PyTorch (3 tiny ops); the decoder receives the
    sampled (non-leaf) latent tensor and returns its gradient."""

    def init_latent_codes(self, dataset_size, ndims, fixed_decoder=True):
        self.log_var = nn.Parameter(torch.normal(-5, 1, size=(dataset_size, ndims, 3)))
        if fixed_decoder:
            self.mu = nn.Parameter(torch.zeros(dataset_size, ndims, 3))
            self.log_var.requires_grad = False
        else:
            self.mu = nn.Parameter(torch.randn((dataset_size, ndims, 3)))

    def sample_latent(self, idx):
        mu = self.mu[idx, :, :]
        log_var = self.log_var[idx, :, :]
        std = torch.exp(0.5 * log_var)
        eps = torch.randn_like(std)
        sample = mu + (eps * std)
        return sample, mu, log_var

    def _latents_for(self, idx):
        if self.fixed_decoder:
            return self.mu[idx, :, :]
        Z, _, _ = self.sample_latent(idx)
        return Z


def get_model(config, dataset_size, task):
    """Reference factory (RENI.py:861-933).

    ``config`` is the reference's yacs node (``config.RENI.*``) or any object with the same attributes."""
    from .film import RENIAutoDecoderFiLM, RENIVariationalAutoDecoderFiLM

    r = config.RENI
    fixed_decoder = task in ["FIT_LATENT", "FIT_INVERSE"]
    if r.CONDITIONING == "Cond-by-Concat":
        cls = {"AutoDecoder": RENIAutoDecoder, "VariationalAutoDecoder": RENIVariationalAutoDecoder}[r.MODEL_TYPE]
        return cls(dataset_size, r.LATENT_DIMENSION, r.EQUIVARIANCE, r.HIDDEN_FEATURES, r.HIDDEN_LAYERS,
                   r.OUT_FEATURES, r.LAST_LAYER_LINEAR, r.OUTPUT_ACTIVATION, r.FIRST_OMEGA_0, r.HIDDEN_OMEGA_0,
                   fixed_decoder)
    if r.CONDITIONING == "FiLM":
        cls = {"AutoDecoder": RENIAutoDecoderFiLM, "VariationalAutoDecoder": RENIVariationalAutoDecoderFiLM}[r.MODEL_TYPE]
        return cls(dataset_size, r.LATENT_DIMENSION, r.EQUIVARIANCE, r.HIDDEN_FEATURES, r.HIDDEN_LAYERS,
                   r.MAPPING_FEATURES, r.MAPPING_LAYERS, r.OUT_FEATURES, r.OUTPUT_ACTIVATION, fixed_decoder)
    return None  # (the reference falls through and returns None for an unknown conditioning)
