"""Drop-in FiLM-conditioned RENI decoders backed by the sm_100a kernels.

Same class names, constructor arguments, attributes, parameter names (``net.{i}.layer.*``, ``final_layer.*``,
``mapping_network.network.{2i}.*``, ``Z`` / ``mu`` / ``log_var``), ``forward`` dispatch and ``load_state_dict``
semantics as the reference's ``RENIAutoDecoderFiLM`` / ``RENIVariationalAutoDecoderFiLM``
(src/models/RENI.py:527-858), which is the reference's default conditioning (configs/default.py:9).

Split of the work (reference ``forward``: RENI.py:653-678):

* per MAP, in PyTorch with autograd (a few (B, .) tensors): the invariant mapping-network input
  (``G = Z_xz Z_xz^T`` and ``Z_y``, RENI.py:418-436), the mapping network (RENI.py:481-512; the reference evaluates
  it on an input replicated over all P pixels -- it is constant per map), ``freq = 15 raw + 30`` (RENI.py:667), and the
  hoisted first FiLM layer: its input ``[|d_xz|, d_y, D_xz Z_xz^T]`` is linear in the direction features, so
  ``freq_0 (W_0 x + b_0) + phase_0 = [f | 1] . mc`` with a per-map (5, 256) matrix ``mc``;
* per DIRECTION, in libreni_b200.so (``reni_film_forward`` / ``reni_film_backward``): the direction features, layer 0,
  the modulated 256x256 layers on the tensor cores, the output layer, and in the backward the gradients w.r.t. ``mc``,
  ``film`` and the decoder weights.  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn

from . import functional as F_
from .functional import FilmSpec, Workspace


def kaiming_leaky_init(m):
    """RENI.py:455-460."""
    if m.__class__.__name__.find("Linear") != -1:
        torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")


class CustomMappingNetwork(nn.Module):
    """RENI.py:481-512: ``map_hidden_layers`` x (Linear, LeakyReLU(0.2)) + Linear; the last weight is scaled by 0.25.
    Evaluated in PyTorch once per map."""

    def __init__(self, in_features, map_hidden_layers, map_hidden_dim, map_output_dim):
        super().__init__()
        network: List[nn.Module] = []
        for _ in range(map_hidden_layers):
            network.append(nn.Linear(in_features, map_hidden_dim))
            network.append(nn.LeakyReLU(0.2, inplace=True))
            in_features = map_hidden_dim
        network.append(nn.Linear(map_hidden_dim, map_output_dim))
        self.network = nn.Sequential(*network)
        self.network.apply(kaiming_leaky_init)
        with torch.no_grad():
            self.network[-1].weight *= 0.25

    def forward(self, z):
        frequencies_offsets = self.network(z)
        half = frequencies_offsets.shape[-1] // 2
        return frequencies_offsets[..., :half], frequencies_offsets[..., half:]


class FiLMLayer(nn.Module):
    """Parameter holder (RENI.py:515-524).  ``forward`` exists for API compatibility; the decoder evaluates
    sin(freq * layer(x) + phase_shift) inside the fused kernel."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.layer = nn.Linear(input_dim, hidden_dim)

    def forward(self, x, freq, phase_shift):  # pragma: no cover - not on the product path
        raise RuntimeError("FiLMLayer is evaluated inside the fused reni_b200 decoder kernel; call the decoder module")


def _frequency_init(layer: nn.Linear, freq: float) -> None:
    """RENI.py:463-471."""
    with torch.no_grad():
        n = layer.weight.size(-1)
        b = math.sqrt(6 / n) / freq
        layer.weight.uniform_(-b, b)


class _FilmMapStage(torch.autograd.Function):
    """(mc, film) = per-map stage(Z; net[0].layer, mapping network) through libreni_b200.so in BOTH directions
    (reni_film_map_forward_train / reni_film_map_backward): what ``_FilmDecoderBase.map_level`` computes with ~40 torch
    ops and differentiates with ~80 more, in 6 + 11 launches.  ``sinks`` (optional) = tensors the parameter gradients
    are accumulated into directly ([dW0, db0, dW_map0, db_map0, ...], e.g. views of the trainer's flat all-reduce
    buffer); the backward then returns None for the parameters.  Reference: RENI.py:405-452, :481-512, :666-678."""

    @staticmethod
    def forward(ctx, cfg, sinks, Z, W0, b0, *map_params):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        B = Z.shape[0]
        H = 256
        L = cfg.hidden_layers
        ws = [p.detach().float().contiguous() for p in map_params[0::2]]
        bs = [p.detach().float().contiguous() for p in map_params[1::2]]
        n = len(ws)
        dims = (C.c_int32 * (n + 1))(ws[0].shape[1], *[w.shape[0] for w in ws])
        Zc = Z.detach().float().contiguous()
        W0c, b0c = W0.detach().float().contiguous(), b0.detach().float().contiguous()
        mc = torch.empty(B, 5, H, device=Z.device, dtype=torch.float32)
        film = torch.empty(B, L, 2, H, device=Z.device, dtype=torch.float32)
        nbytes = int(lib.reni_film_map_acts_bytes(dims, n, B))
        if nbytes < 0:
            _lib.check(nbytes, "reni_film_map_acts_bytes")
        acts = torch.empty(nbytes, dtype=torch.uint8, device=Z.device)
        ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        stream = C.c_void_p(torch.cuda.current_stream(Z.device).cuda_stream)
        rc = lib.reni_film_map_forward_train(C.byref(cfg), C.c_void_p(Zc.data_ptr()), C.c_void_p(W0c.data_ptr()),
                                             C.c_void_p(b0c.data_ptr()), ptrs(ws), ptrs(bs), dims, n, B,
                                             C.c_void_p(mc.data_ptr()), C.c_void_p(film.data_ptr()),
                                             C.c_void_p(acts.data_ptr()), nbytes, stream)
        _lib.check(rc, "reni_film_map_forward_train")
        ctx.cfg, ctx.sinks, ctx.dims, ctx.n, ctx.nbytes = cfg, sinks, dims, n, nbytes
        ctx.save_for_backward(Zc, W0c, b0c, acts, *ws)
        return mc, film

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_mc, d_film):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        Zc, W0c, b0c, acts = ctx.saved_tensors[:4]
        ws = list(ctx.saved_tensors[4:])
        n, B, dev = ctx.n, Zc.shape[0], Zc.device
        L = ctx.cfg.hidden_layers
        d_mc = torch.zeros(B, 5, 256, device=dev) if d_mc is None else d_mc.float().contiguous()
        d_film = torch.zeros(B, L, 2, 256, device=dev) if d_film is None else d_film.float().contiguous()
        want_dw = any(ctx.needs_input_grad[3:])
        sinks = ctx.sinks
        own = None
        if want_dw and sinks is None:
            own = [torch.zeros_like(W0c), torch.zeros_like(b0c)]
            for w in ws:
                own += [torch.zeros_like(w), torch.zeros(w.shape[0], device=dev, dtype=torch.float32)]
            sinks = own
        dZ = torch.empty_like(Zc)
        scratch_bytes = ctx.nbytes + B * 4 * 256 * 4
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
        ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        rc = lib.reni_film_map_backward(
            C.byref(ctx.cfg), vp(Zc), vp(W0c), vp(b0c), ptrs(ws), ctx.dims, n, B, vp(acts), vp(d_mc), vp(d_film), vp(dZ),
            vp(sinks[0]) if want_dw else None, vp(sinks[1]) if want_dw else None,
            ptrs(sinks[2::2]) if want_dw else None, ptrs(sinks[3::2]) if want_dw else None, vp(scratch), scratch_bytes,
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc, "reni_film_map_backward")
        grads = [None] * (2 + 2 * n)
        if own is not None:
            grads = [g if need else None for g, need in zip(own, ctx.needs_input_grad[3:])]
        return (None, None, dZ if ctx.needs_input_grad[2] else None, *grads)


class _FilmDecoderBase(nn.Module):
    def __init__(self, dataset_size, ndims, equivariance, siren_hidden_features, siren_hidden_layers,
                 mapping_network_features, mapping_network_layers, out_features, output_activation, fixed_decoder):
        super().__init__()
        self.dataset_size = dataset_size
        self.ndims = ndims
        self.equivariance = equivariance
        self.siren_hidden_features = siren_hidden_features
        self.siren_hidden_layers = siren_hidden_layers
        self.mapping_network_features = mapping_network_features
        self.mapping_network_layers = mapping_network_layers
        self.out_features = out_features
        self.output_activation = output_activation
        self.fixed_decoder = fixed_decoder

        if equivariance == "None":     # RENI.py:556-567
            self.in_features = ndims * 3
            self.mn_in_features = ndims
        elif equivariance == "SO2":
            self.in_features = 2 + ndims
            self.mn_in_features = ndims * ndims + ndims
        elif equivariance == "SO3":
            self.in_features = ndims
            self.mn_in_features = ndims * ndims
        else:
            raise ValueError(f"unknown equivariance {equivariance!r}")

        self.init_latent_codes(dataset_size, ndims, fixed_decoder)

        self.net = nn.ModuleList()
        self.net.append(FiLMLayer(self.in_features, siren_hidden_features))
        for _ in range(siren_hidden_layers - 1):
            self.net.append(FiLMLayer(siren_hidden_features, siren_hidden_features))
        self.final_layer = nn.Linear(siren_hidden_features, out_features)
        self.mapping_network = CustomMappingNetwork(self.mn_in_features, mapping_network_layers,
                                                    mapping_network_features,
                                                    len(self.net) * siren_hidden_features * 2)
        for layer in self.net:         # RENI.py:584-586
            _frequency_init(layer.layer, 25)
        _frequency_init(self.final_layer, 25)
        with torch.no_grad():
            n = self.net[0].layer.weight.size(-1)
            self.net[0].layer.weight.uniform_(-1 / n, 1 / n)

        if fixed_decoder:              # RENI.py:595-601
            for group in (self.net, self.final_layer, self.mapping_network):
                for param in group.parameters():
                    param.requires_grad = False
        self._ws = Workspace()

    # ---- plumbing -------------------------------------------------------------------------
    @property
    def spec(self) -> FilmSpec:
        return FilmSpec(self.ndims, self.equivariance, self.siren_hidden_features, self.siren_hidden_layers,
                        self.out_features, self.output_activation)

    def core_parameters(self) -> List[torch.Tensor]:
        """[W_1, b_1, ..., W_L, b_L, W_out, b_out]: what the fused kernel streams (net.0 is hoisted per map)."""
        ps: List[torch.Tensor] = []
        for layer in list(self.net)[1:]:
            ps += [layer.layer.weight, layer.layer.bias]
        return ps + [self.final_layer.weight, self.final_layer.bias]

    # No-grad decodes of up to this many latents take the native per-map stage (reni_film_map_forward: 6 launches
    # instead of ~25 torch ops; it re-reads the mapping-network weights from L2 for every map, so large batches keep
    # the batched torch GEMMs).  The two stages agree to fp32 rounding (2e-6), not bit for bit.
    NATIVE_MAP_LEVEL_MAX_BATCH = 128

    def _map_level_native(self, Z: torch.Tensor):
        """The same per-map operands from reni_film_map_forward (2 + n_linears launches of GEMV-style kernels instead
        of ~25 torch launches): for no-grad decoding of a few latents; every map re-reads the weights from L2, hence
        the cap on the batch."""
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        B = Z.shape[0]
        H, L = self.siren_hidden_features, self.siren_hidden_layers - 1
        Zc = Z.detach().float().contiguous()
        lins = [m for m in self.mapping_network.network if isinstance(m, nn.Linear)]
        ws = [m.weight.detach().float().contiguous() for m in lins]
        bs = [m.bias.detach().float().contiguous() for m in lins]
        dims = (C.c_int32 * (len(lins) + 1))(lins[0].in_features, *[m.out_features for m in lins])
        W0 = self.net[0].layer.weight.detach().float().contiguous()
        b0 = self.net[0].layer.bias.detach().float().contiguous()
        mc = torch.empty(B, 5, H, device=Z.device, dtype=torch.float32)
        film = torch.empty(B, L, 2, H, device=Z.device, dtype=torch.float32)
        ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        nbytes = int(lib.reni_film_map_scratch_bytes(dims, len(lins), B))
        if nbytes < 0:
            _lib.check(nbytes, "reni_film_map_scratch_bytes")
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=Z.device)
        rc = lib.reni_film_map_forward(
            C.byref(self.spec.c_config()), C.c_void_p(Zc.data_ptr()), C.c_void_p(W0.data_ptr()),
            C.c_void_p(b0.data_ptr()), ptrs(ws), ptrs(bs), dims, len(lins), B, C.c_void_p(mc.data_ptr()),
            C.c_void_p(film.data_ptr()), C.c_void_p(scratch.data_ptr()), nbytes,
            C.c_void_p(torch.cuda.current_stream(Z.device).cuda_stream))
        _lib.check(rc, "reni_film_map_forward")
        return mc, film

    def _map_params(self) -> List[torch.Tensor]:
        """[W0, b0, W_map0, b_map0, ...]: the parameters of the per-map stage, in the order of its gradient sinks."""
        ps: List[torch.Tensor] = [self.net[0].layer.weight, self.net[0].layer.bias]
        for m in self.mapping_network.network:
            if isinstance(m, nn.Linear):
                ps += [m.weight, m.bias]
        return ps

    def map_level(self, Z: torch.Tensor, grad_sinks=None):
        """Per-map operands of the core: mc (B, 5, H) and film (B, L, 2, H), differentiable w.r.t. Z, net[0] and the
        mapping network.  On the GPU, for up to NATIVE_MAP_LEVEL_MAX_BATCH maps, the stage runs natively in both
        directions (``_FilmMapStage``; RENI_FILM_NATIVE_MAP=0 keeps the torch ops below, which are also what larger
        batches use).  ``grad_sinks``: see ``_FilmMapStage``."""
        import os

        n_lin = len([m for m in self.mapping_network.network if isinstance(m, nn.Linear)])
        if (Z.is_cuda and Z.dtype == torch.float32 and Z.shape[0] <= self.NATIVE_MAP_LEVEL_MAX_BATCH and n_lin <= 8
                and self.equivariance in ("SO2", "SO3") and os.environ.get("RENI_FILM_NATIVE_MAP", "1") != "0"
                and self.net[0].layer.weight.dtype == torch.float32):
            return _FilmMapStage.apply(self.spec.c_config(), grad_sinks, Z, *self._map_params())
        B, N, _ = Z.shape
        H, Lf = self.siren_hidden_features, self.siren_hidden_layers
        W0, b0 = self.net[0].layer.weight, self.net[0].layer.bias
        if self.equivariance == "SO2":
            Z_xz = torch.stack((Z[:, :, 0], Z[:, :, 2]), -1)
            G = torch.bmm(Z_xz, Z_xz.transpose(1, 2))
            mapping_input = torch.cat((G.flatten(start_dim=1), Z[:, :, 1]), 1)        # RENI.py:424-435
            W_ip = W0[:, 2:]
            M = torch.stack((Z_xz[:, :, 0] @ W_ip.T, Z_xz[:, :, 1] @ W_ip.T,          # d_x, d_z
                             W0[:, 0].expand(B, H), W0[:, 1].expand(B, H)), 1)        # |d_xz|, d_y (RENI.py:434)
        else:  # SO3 (RENI.py:405-415)
            G = Z @ Z.transpose(1, 2)
            mapping_input = G.flatten(start_dim=1)
            M = torch.cat((Z.transpose(1, 2) @ W0.T, torch.zeros(B, 1, H, device=Z.device, dtype=Z.dtype)), 1)
        frequencies, phase_shifts = self.mapping_network(mapping_input)
        frequencies = frequencies * 15 + 30                                             # RENI.py:667
        freq = frequencies.view(B, Lf, H)
        phase = phase_shifts.view(B, Lf, H)
        mc = torch.cat((M * freq[:, :1], (freq[:, 0] * b0 + phase[:, 0]).unsqueeze(1)), 1)
        film = torch.stack((freq[:, 1:], phase[:, 1:]), 2)
        return mc, film

    def decode(self, Z: torch.Tensor, directions: torch.Tensor) -> torch.Tensor:
        spec = self.spec
        spec.validate()
        if Z.dim() != 3 or Z.shape[1] != self.ndims or Z.shape[2] != 3:
            raise ValueError(f"latent codes must have shape (B, {self.ndims}, 3), got {tuple(Z.shape)}")
        if Z.device.type != "cuda":
            raise RuntimeError("reni_b200 runs on CUDA (sm_100a) only and has no CPU fallback; got latent codes on "
                               f"{Z.device}.  Move the module and its inputs to a B200.")
        if Z.shape[0] == 0 or (directions.dim() == 3 and directions.shape[1] == 0):
            out = F_.empty_output(Z, directions, [p for n, p in self.named_parameters() if n not in ("Z", "mu", "log_var")],
                                  self.out_features)
            return torch.exp(out) if self.output_activation == "exp" else out
        differentiated = torch.is_grad_enabled() and (Z.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not differentiated and Z.shape[0] <= self.NATIVE_MAP_LEVEL_MAX_BATCH and len(
                [m for m in self.mapping_network.network if isinstance(m, nn.Linear)]) <= 8:
            mc, film = self._map_level_native(Z)
        else:
            mc, film = self.map_level(Z.float())
        out = F_.film_decode_core(spec, self._ws, mc, film, directions, self.core_parameters())
        if self.output_activation == "exp":
            out = torch.exp(out)
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        """Reference semantics (RENI.py:608-627): keep keys prefixed ``model.``; a fixed decoder loads net,
        mapping_network and final_layer only and keeps its fresh latents."""
        new_state_dict = {k[6:]: v for k, v in state_dict.items() if k.startswith("model.")}
        if self.fixed_decoder:
            net_sd = {k[4:]: v for k, v in new_state_dict.items() if k.startswith("net.")}
            map_sd = {k[16:]: v for k, v in new_state_dict.items() if k.startswith("mapping_network.")}
            self.net.load_state_dict(net_sd, strict=strict)
            self.mapping_network.load_state_dict(map_sd, strict=strict)
            dev = self.final_layer.weight.device
            self.final_layer.weight = nn.Parameter(new_state_dict["final_layer.weight"].to(dev), requires_grad=False)
            self.final_layer.bias = nn.Parameter(new_state_dict["final_layer.bias"].to(dev), requires_grad=False)
            return None
        return super().load_state_dict(new_state_dict, strict=strict)

    # ---- forward dispatch (RENI.py:629-664 / :806-858) ----------------------------------------
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


class RENIAutoDecoderFiLM(_FilmDecoderBase):
    """Reference: src/models/RENI.py:527-678."""

    def init_latent_codes(self, dataset_size, ndims, fixed_decoder=False):
        if fixed_decoder:
            self.Z = nn.Parameter(torch.zeros(dataset_size, ndims, 3))
        else:
            self.Z = nn.Parameter(torch.randn((dataset_size, ndims, 3)))

    def _latents_for(self, idx):
        return self.Z[idx, :, :]


class RENIVariationalAutoDecoderFiLM(_FilmDecoderBase):
    """Reference: src/models/RENI.py:681-858.  Sampling stays in PyTorch; the decoder differentiates through the
    sampled (non-leaf) latent tensor."""

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
