"""Host-side mirror of the decoder call: torch tensors in, C-ABI kernels underneath.

``decode`` is the autograd-aware replacement for ``self.net(self.InvariantRepresentation(Z, D))``
(reference: src/models/RENI.py:211-233); ``loss_forward_backward`` is the fused
``training_step`` body (src/lightning/RENI_module.py:80-146 with RENITrainLoss / RENITestLoss,
src/utils/loss_functions.py:39-71).  PyTorch is used for device memory and streams only; all
arithmetic happens in libreni_b200.so.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FLAG_LOSS, FLAG_NEED_DW, FLAG_SAVE_FOR_BACKWARD, RENIConfig

HIDDEN_FEATURES = 256  # compile-time tile width of the tcgen05 kernels


@dataclass(frozen=True)
class DecoderSpec:
    """Hyper-parameters of one decoder, as the reference constructor takes them (RENI.py:91-104)."""

    ndims: int
    equivariance: str
    hidden_features: int
    hidden_layers: int
    out_features: int
    last_layer_linear: bool
    output_activation: Optional[str]
    first_omega_0: float
    hidden_omega_0: float

    def validate(self) -> None:
        if self.equivariance not in _lib.EQUIVARIANCE:
            raise ValueError(f"equivariance must be one of {list(_lib.EQUIVARIANCE)}, got {self.equivariance!r}")
        if self.output_activation == "exp":
            # the reference raises here too: nn.Exp does not exist (RENI.py:173-174)
            raise AttributeError("module 'torch.nn' has no attribute 'Exp'")
        if self.output_activation not in (None, "tanh"):
            raise ValueError(f"unsupported output_activation {self.output_activation!r}")
        if not 1 <= self.hidden_features <= HIDDEN_FEATURES:
            raise NotImplementedError(
                f"reni_b200 kernels are built for hidden_features <= {HIDDEN_FEATURES} (narrower decoders run zero-padded "
                f"to {HIDDEN_FEATURES}), got {self.hidden_features}")
        if not 1 <= self.hidden_layers <= 6:
            raise NotImplementedError("reni_b200 kernels support 1..6 hidden layers")
        if not 1 <= self.out_features <= 3:
            raise NotImplementedError("reni_b200 kernels support out_features <= 3")

    def padded(self) -> "DecoderSpec":
        """The spec the kernels run: hidden width 256.  A narrower decoder is embedded exactly -- zero rows / columns give
        sin(0) = 0 activations that feed nothing, and zero gradients outside the real block (``pad_parameters``)."""
        if self.hidden_features == HIDDEN_FEATURES:
            return self
        import dataclasses

        return dataclasses.replace(self, hidden_features=HIDDEN_FEATURES)

    def c_config(self) -> RENIConfig:
        self.validate()
        if self.hidden_features != HIDDEN_FEATURES:
            return self.padded().c_config()
        return RENIConfig(
            self.ndims, _lib.EQUIVARIANCE[self.equivariance], self.hidden_features, self.hidden_layers,
            self.out_features, 1 if self.last_layer_linear else 0, 1 if self.output_activation == "tanh" else 0,
            float(self.first_omega_0), float(self.hidden_omega_0))


def pad_parameters(spec: DecoderSpec, params: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """[W0, b0, ..., W_out, b_out] of a decoder with hidden width H < 256 zero-padded to the kernels' width (differentiable:
    ``F.pad``'s backward slices the gradients back).  Identity at H = 256."""
    H = spec.hidden_features
    if H == HIDDEN_FEATURES:
        return list(params)
    pad = HIDDEN_FEATURES - H
    n = len(params) // 2
    out: List[torch.Tensor] = []
    for i in range(n):
        W, b = params[2 * i], params[2 * i + 1]
        last = i == n - 1
        if i == 0:
            Wp = torch.nn.functional.pad(W, (0, 0, 0, pad))            # (H, in) -> (256, in)
        elif last:
            Wp = torch.nn.functional.pad(W, (0, pad))                  # (out, H) -> (out, 256)
        else:
            Wp = torch.nn.functional.pad(W, (0, pad, 0, pad))          # (H, H) -> (256, 256)
        bp = b if last else torch.nn.functional.pad(b, (0, pad))
        out += [Wp, bp]
    return out


def _vp(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    return (C.c_void_p * len(ts))(*[(t.data_ptr() if t is not None else 0) for t in ts])


def _call(dev: torch.device, fn, *args):
    """Library call with ``dev`` as the current CUDA device: the library sizes its grids, and creates its side stream,
    for the current device, while the kernels run on the tensors' stream -- the two must agree (a model on cuda:1
    with cuda:0 current would otherwise fork onto a stream of the wrong device)."""
    if dev.index is not None and torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            return fn(*args)
    return fn(*args)


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError(
                "reni_b200 runs on CUDA (sm_100a) only and has no CPU fallback; got a tensor on "
                f"{t.device}.  Move the module and its inputs to a B200.")
        if t.device != dev:
            raise RuntimeError(f"all tensors must be on the same device ({dev} vs {t.device})")
    return dev


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _batch_stride(t: torch.Tensor, B: int, what: str) -> Tuple[torch.Tensor, int]:
    """(tensor to pass, elements between maps).  A (1,P,3) tensor or a stride-0 expand is shared by all maps."""
    if t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"{what} must have shape (B, P, 3), got {tuple(t.shape)}")
    if t.shape[0] == 1 or (t.stride(0) == 0 and t.shape[0] == B):
        t0 = _f32c(t[:1])
        return t0, 0
    # the reference asserts len(idx) == directions.shape[0] (RENI.py:213,220)
    assert t.shape[0] == B, f"{what} batch {t.shape[0]} != latent batch {B}"
    t = _f32c(t)
    return t, t.shape[1] * 3


class Workspace:
    """Caller-owned device workspace (grow-only), 1024-byte aligned, carved by the library."""

    def __init__(self) -> None:
        self.buf: Optional[torch.Tensor] = None
        self.view: Optional[torch.Tensor] = None
        self.nbytes = 0
        self.prepared_key = None

    def ensure(self, nbytes: int, device: torch.device) -> torch.Tensor:
        if self.view is None or self.nbytes < nbytes or self.view.device != device:
            self.buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-self.buf.data_ptr()) % 1024
            self.view = self.buf[off:off + nbytes]
            self.nbytes = nbytes
            self.prepared_key = None
        return self.view


def workspace_bytes(cfg: RENIConfig, B: int, P: int, flags: int) -> int:
    n = _lib.load().reni_workspace_bytes(C.byref(cfg), B, P, flags)
    if n <= 0:
        _lib.check(int(n), "reni_workspace_bytes")
    return int(n)


def _bwd_schedule_flag(tile_major: Optional[bool]) -> int:
    """RENI_FLAG_TILE_MAJOR_BWD / RENI_FLAG_LAYER_MAJOR_BWD from an explicit choice or the RENI_TILE_MAJOR_BWD variable."""
    if tile_major is None:
        env = os.environ.get("RENI_TILE_MAJOR_BWD")
        if env is None:
            return 0
        tile_major = env == "1"
    return _lib.FLAG_TILE_MAJOR_BWD if tile_major else _lib.FLAG_LAYER_MAJOR_BWD


def _fwd_terms_flag() -> int:
    """RENI_FWD_TERMS=1 / 2 forces one- / two-term forward weights (A/B switch; default: the library's, two-term)."""
    env = os.environ.get("RENI_FWD_TERMS")
    if env == "1":
        return _lib.FLAG_FWD_SINGLE_TERM
    if env == "2":
        return _lib.FLAG_FWD_TWO_TERM
    return 0


def _params_key(weights, biases):
    return tuple((t.data_ptr(), t._version) for t in list(weights) + list(biases))


def prepare_weights(cfg: RENIConfig, weights, biases, ws: Workspace, device) -> None:
    """fp32 parameters -> fp16 operand images inside the workspace (skipped if unchanged)."""
    dev = device
    key = _params_key(weights, biases)
    if ws.prepared_key == key:
        return
    lib = _lib.load()
    rc = _call(dev, lib.reni_prepare_weights, C.byref(cfg), _ptr_array(weights), _ptr_array(biases), _vp(ws.view), ws.nbytes,
                                  _stream(device))
    _lib.check(rc, "reni_prepare_weights")
    ws.prepared_key = key


class _DecodeFunction(torch.autograd.Function):
    """out = net(encoding(Z, D)) with gradients for Z and (optionally) the decoder parameters."""

    @staticmethod
    def forward(ctx, spec: DecoderSpec, inference_ws: Workspace, Z, D, *params):
        lib = _lib.load()
        cfg = spec.c_config()
        nl = spec.hidden_layers + 2
        weights = [_f32c(p) for p in params[0::2]]
        biases = [_f32c(p) for p in params[1::2]]
        assert len(weights) == nl and len(biases) == nl
        dev = _require_cuda(Z, D, *weights, *biases)
        if Z.dim() != 3 or Z.shape[1] != spec.ndims or Z.shape[2] != 3:
            raise ValueError(f"latent codes must have shape (B, {spec.ndims}, 3), got {tuple(Z.shape)}")
        Zc = _f32c(Z)
        B = Zc.shape[0]
        Dc, d_bs = _batch_stride(D, B, "directions")
        P = Dc.shape[1]
        need_dz = ctx.needs_input_grad[2]
        need_dw = any(ctx.needs_input_grad[4:])
        flags = _fwd_terms_flag()
        if need_dz or need_dw:
            flags |= FLAG_SAVE_FOR_BACKWARD
            if need_dw:
                flags |= FLAG_NEED_DW
                flags |= _bwd_schedule_flag(None)
        nbytes = workspace_bytes(cfg, B, P, flags)
        # a forward that will be differentiated owns its stash until backward has run
        ws = Workspace() if (flags & FLAG_SAVE_FOR_BACKWARD) else inference_ws
        ws.ensure(nbytes, dev)
        prepare_weights(cfg, weights, biases, ws, dev)
        out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
        rc = _call(dev, lib.reni_forward, C.byref(cfg), _vp(Zc), _vp(Dc), d_bs, _vp(weights[0]), _vp(biases[0]), B, P, _vp(out),
                              None, None, 0, _vp(ws.view), ws.nbytes, flags, _stream(dev))
        _lib.check(rc, "reni_forward")
        if flags & FLAG_SAVE_FOR_BACKWARD:
            ctx.spec, ctx.ws, ctx.flags, ctx.d_bs = spec, ws, flags, d_bs
            ctx.save_for_backward(Zc, Dc, out, *weights, *biases)
        if spec.out_features != 3:
            return out[:, :, : spec.out_features]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        spec: DecoderSpec = ctx.spec
        cfg = spec.c_config()
        nl = spec.hidden_layers + 2
        saved = ctx.saved_tensors
        Zc, Dc, out = saved[0], saved[1], saved[2]
        weights, biases = list(saved[3:3 + nl]), list(saved[3 + nl:3 + 2 * nl])
        dev = Zc.device
        B, P = Zc.shape[0], Dc.shape[1]
        g = _f32c(grad_out)
        if g.shape[2] != 3:
            gp = torch.zeros(B, P, 3, device=dev)
            gp[:, :, : g.shape[2]] = g
            g = gp
        need_dw = bool(ctx.flags & FLAG_NEED_DW)
        dZ = torch.empty_like(Zc)
        dW = [torch.zeros_like(w) for w in weights] if need_dw else None
        db = [torch.zeros_like(b) for b in biases] if need_dw else None
        rc = _call(dev, lib.reni_backward, C.byref(cfg), _vp(Zc), _vp(Dc), ctx.d_bs, _ptr_array(weights), B, P, _vp(out), _vp(g),
                               _vp(dZ), _ptr_array(dW) if need_dw else None, _ptr_array(db) if need_dw else None,
                               _vp(ctx.ws.view), ctx.ws.nbytes, ctx.flags, _stream(dev))
        _lib.check(rc, "reni_backward")
        ctx.ws = None  # release the stash
        grads: List[Optional[torch.Tensor]] = []
        for i in range(nl):
            grads.append(dW[i] if need_dw and ctx.needs_input_grad[4 + 2 * i] else None)
            grads.append(db[i] if need_dw and ctx.needs_input_grad[5 + 2 * i] else None)
        return (None, None, dZ if ctx.needs_input_grad[2] else None, None, *grads)


def empty_output(Z: torch.Tensor, D: torch.Tensor, params: Sequence[torch.Tensor], out_features: int) -> torch.Tensor:
    """Empty batch (B = 0) or empty direction set (P = 0): the reference's ops are shape-generic and return an empty
    (B, P, out_features) tensor whose backward leaves zero gradients; no kernel runs (the C ABI refuses B < 1, P < 1)."""
    _require_cuda(Z, D, *params)
    out = torch.zeros(Z.shape[0], D.shape[1], out_features, device=Z.device, dtype=torch.float32)
    if torch.is_grad_enabled():
        tie = None
        for t in (Z, *params):
            if t.requires_grad:
                tie = t.sum() * 0 if tie is None else tie + t.sum() * 0
        if tie is not None:
            out = out + tie
    return out


def decode(spec: DecoderSpec, inference_ws: Workspace, Z: torch.Tensor, D: torch.Tensor,
           params: Sequence[torch.Tensor]) -> torch.Tensor:
    """``params`` = [W0, b0, W1, b1, ..., W_out, b_out] (state_dict order of ``net``).

    Differentiable w.r.t. Z and the parameters (never w.r.t. the directions, which the reference never
    differentiates either).  Under ``torch.no_grad()`` or with nothing requiring grad this is the
    inference kernel with no stash."""
    spec.validate()
    if Z.dim() == 3 and D.dim() == 3 and (Z.shape[0] == 0 or D.shape[1] == 0):
        return empty_output(Z, D, params, spec.out_features)
    if torch.is_grad_enabled() and (Z.requires_grad or any(p.requires_grad for p in params)):
        return _DecodeFunction.apply(spec.padded(), inference_ws, Z, D, *pad_parameters(spec, params))
    with torch.no_grad():
        return _DecodeFunction.forward(_NoGradCtx(len(params)), spec.padded(), inference_ws, Z, D,
                                       *pad_parameters(spec, params))


class _NoGradCtx:
    """Stand-in ctx for calling the forward outside autograd."""

    def __init__(self, nparams: int) -> None:
        self.needs_input_grad = (False,) * (4 + nparams)

    def save_for_backward(self, *a) -> None:  # pragma: no cover - never reached (flags == 0)
        raise AssertionError


@dataclass
class StepResult:
    """What the reference's training_step returns (RENI_module.py:115-146) plus the gradients."""

    loss: torch.Tensor         # scalar
    mse_loss: torch.Tensor
    prior_loss: torch.Tensor
    cosine_loss: torch.Tensor
    out: torch.Tensor          # (B, P, 3) model output
    dZ: torch.Tensor           # (B, N, 3)
    dW: Optional[List[torch.Tensor]]
    db: Optional[List[torch.Tensor]]


def loss_forward_backward(spec: DecoderSpec, ws: Workspace, Z: torch.Tensor, D: Optional[torch.Tensor],
                          target: torch.Tensor, sineweight: Optional[torch.Tensor], weights: Sequence[torch.Tensor],
                          biases: Sequence[torch.Tensor],
                          alpha: float = 0.0, beta: float = 0.0, use_cosine: bool = False, need_dw: bool = True,
                          grad_weights: Optional[Sequence[torch.Tensor]] = None,
                          grad_biases: Optional[Sequence[torch.Tensor]] = None,
                          tile_major_bwd: Optional[bool] = None, mask_bits: Optional[torch.Tensor] = None) -> StepResult:
    """Fused forward + loss + backward (one training / latent-fit step without the optimiser).

    Analytic grid (the reference's callers always pass ``get_directions(W)`` / ``get_sineweight(W) [* mask]``,
    RENI_module.py:89-94): ``D=None`` makes the kernels compute the equirectangular directions from the pixel index
    (P = W * W / 2), ``sineweight=None`` the sine weights, times ``mask_bits`` (``geometry.pack_mask_bits``) if given.

    ``tile_major_bwd``: True forces the tile-major delta chain + split-K weight-gradient GEMM, False the layer-major
    backward (one launch per layer); None takes ``RENI_TILE_MAJOR_BWD`` (0/1) or else the library default (tile-major).

    loss = WeightedMSE + alpha * sum Z^2 + beta * WeightedCosineSimilarity  (loss_functions.py:6-32,60-71);
    RENITrainLoss is alpha = beta = 0.  ``grad_weights`` / ``grad_biases`` (e.g. views into one flat
    all-reduce buffer) are ACCUMULATED into; fresh zero tensors are used when omitted."""
    if spec.hidden_features != HIDDEN_FEATURES:
        return _narrow_loss_forward_backward(spec, ws, Z, D, target, sineweight, weights, biases, alpha, beta, use_cosine,
                                             need_dw, grad_weights, grad_biases, tile_major_bwd, mask_bits)
    lib = _lib.load()
    cfg = spec.c_config()
    dev = _require_cuda(Z, target, *[t for t in (D, sineweight, mask_bits) if t is not None], *weights, *biases)
    Zc = _f32c(Z)
    B = Zc.shape[0]
    P = target.shape[1]
    grid_flags = 0
    if D is None:
        Dc, d_bs = None, 0
        grid_flags |= _lib.FLAG_GRID_DIRECTIONS
    else:
        Dc, d_bs = _batch_stride(D, B, "directions")
        if Dc.shape[1] != P:
            raise ValueError(f"directions have {Dc.shape[1]} pixels, target {P}")
    if sineweight is None:
        swc, sw_bs = None, 0
        grid_flags |= _lib.FLAG_GRID_SINEWEIGHT
        if mask_bits is not None:
            if mask_bits.dtype != torch.int32 or mask_bits.numel() * 32 < P:
                raise ValueError("mask_bits must be int32 words with one bit per pixel (geometry.pack_mask_bits)")
            swc = mask_bits.contiguous()
    else:
        if mask_bits is not None:
            raise ValueError("mask_bits goes with sineweight=None (analytic sine weights); multiply a sineweight tensor by the mask instead")
        swc, sw_bs = _batch_stride(sineweight, B, "sineweight")
    if grid_flags:
        side = int(round((2 * P) ** 0.5))
        if side * side // 2 != P or side % 2:
            raise ValueError(f"the analytic grid needs P = W * W / 2 directions, got {P}")
    tc = _f32c(target)
    if tuple(tc.shape) != (B, P, 3):
        raise ValueError(f"target must have shape {(B, P, 3)}, got {tuple(tc.shape)}")
    weights = [_f32c(w) for w in weights]
    biases = [_f32c(b) for b in biases]
    flags = FLAG_SAVE_FOR_BACKWARD | FLAG_LOSS | (FLAG_NEED_DW if need_dw else 0) | grid_flags
    flags |= _bwd_schedule_flag(tile_major_bwd) | _fwd_terms_flag()
    # Maps are independent units and every gradient buffer is accumulated into, so a batch whose stash would not fit the
    # workspace budget is walked in chunks of maps through the same workspace (BASELINE configs[3] at full size: 4096
    # maps would need 97 GiB in one call).  RENI_MAX_WORKSPACE_GB (default 48) sets the budget.
    budget = int(float(os.environ.get("RENI_MAX_WORKSPACE_GB", "48")) * 2 ** 30)
    chunk = B
    if workspace_bytes(cfg, B, P, flags) > budget and B > 1:
        per_map = max(1, (workspace_bytes(cfg, min(B, 64), P, flags) - workspace_bytes(cfg, 1, P, flags)) // max(1, min(B, 64) - 1))
        chunk = max(1, min(B, (budget - workspace_bytes(cfg, 1, P, flags)) // per_map))
    ws.ensure(workspace_bytes(cfg, chunk, P, flags), dev)
    # changed parameters (every training step): the library rebuilds the fp16 weight images inside the fused call, on
    # its side stream beside the per-map prologue, instead of in a launch of its own ahead of it
    key = _params_key(weights, biases)
    if ws.prepared_key != key:
        if os.environ.get("RENI_PREPARE_IN_CALL", "0") == "1" and chunk == B:
            flags |= _lib.FLAG_PREPARE_WEIGHTS
        else:
            prepare_weights(cfg, weights, biases, ws, dev)
    out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
    loss = torch.empty(4, device=dev, dtype=torch.float32)
    dZ = torch.empty_like(Zc)
    dW = db = None
    if need_dw:
        dW = list(grad_weights) if grad_weights is not None else [torch.zeros_like(w) for w in weights]
        db = list(grad_biases) if grad_biases is not None else [torch.zeros_like(b) for b in biases]
    total = None
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        part = loss if chunk == B else torch.empty(4, device=dev, dtype=torch.float32)
        rc = _call(dev, lib.reni_loss_forward_backward,
                   C.byref(cfg), _vp(Zc[lo:hi]), _vp(Dc if d_bs == 0 or Dc is None else Dc[lo:hi]), d_bs,
                   _ptr_array(weights), _ptr_array(biases), hi - lo, P, _vp(tc[lo:hi]),
                   _vp(swc if sw_bs == 0 or swc is None else swc[lo:hi]), sw_bs,
                   float(alpha), float(beta), 1 if use_cosine else 0, _vp(out[lo:hi]), _vp(part), _vp(dZ[lo:hi]),
                   _ptr_array(dW) if need_dw else None, _ptr_array(db) if need_dw else None, _vp(ws.view), ws.nbytes,
                   flags, _stream(dev))
        _lib.check(rc, "reni_loss_forward_backward")
        if chunk != B:
            total = part if total is None else total + part
    if total is not None:
        loss = total
    ws.prepared_key = key
    return StepResult(loss[0], loss[1], loss[2], loss[3], out, dZ, dW, db)


def _narrow_loss_forward_backward(spec, ws, Z, D, target, sineweight, weights, biases, alpha, beta, use_cosine, need_dw,
                                  grad_weights, grad_biases, tile_major_bwd, mask_bits) -> StepResult:
    """Fused step of a decoder narrower than the kernels' 256 features: run the zero-padded decoder, hand back the real
    blocks of the gradients (accumulated into ``grad_weights`` / ``grad_biases`` like the full-width path)."""
    with torch.no_grad():
        flat = []
        for wt, bt in zip(weights, biases):
            flat += [wt, bt]
        padded = pad_parameters(spec, flat)
    r = loss_forward_backward(spec.padded(), ws, Z, D, target, sineweight, padded[0::2], padded[1::2], alpha=alpha,
                              beta=beta, use_cosine=use_cosine, need_dw=need_dw, tile_major_bwd=tile_major_bwd,
                              mask_bits=mask_bits)
    if not need_dw:
        return r
    dW, db = [], []
    for i, (wt, bt) in enumerate(zip(weights, biases)):
        gw = r.dW[i][: wt.shape[0], : wt.shape[1]]
        gb = r.db[i][: bt.shape[0]]
        if grad_weights is not None:
            grad_weights[i].add_(gw)
            grad_biases[i].add_(gb)
            dW.append(grad_weights[i])
            db.append(grad_biases[i])
        else:
            dW.append(gw.contiguous())
            db.append(gb.contiguous())
    return StepResult(r.loss, r.mse_loss, r.prior_loss, r.cosine_loss, r.out, r.dZ, dW, db)


# ----------------------------------------------------------------------------------------------------------------
# FiLM-conditioned decoder core (reference: RENIAutoDecoderFiLM.forward_with_frequencies_phase_shifts,
# src/models/RENI.py:666-678 with FiLMLayer :515-524)
# ----------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FilmSpec:
    """Hyper-parameters of one FiLM decoder as the reference constructor takes them (RENI.py:528-540)."""

    ndims: int
    equivariance: str
    siren_hidden_features: int
    siren_hidden_layers: int
    out_features: int
    output_activation: Optional[str]

    def validate(self) -> None:
        if self.equivariance not in ("SO2", "SO3"):
            # the reference builds net[0] = Linear(3N, H) for "None" but feeds it the (B, P, N) inner products and
            # gives the mapping network N instead of 3N inputs (RENI.py:556-559,438-441): its forward raises
            raise RuntimeError("equivariance 'None' with FiLM conditioning is broken in the reference (shape mismatch "
                               "in net[0] and the mapping network, RENI.py:438-441,556-559)")
        if self.output_activation not in (None, "tanh", "exp"):
            raise ValueError(f"unsupported output_activation {self.output_activation!r}")
        if self.siren_hidden_features != HIDDEN_FEATURES:
            raise NotImplementedError(
                f"reni_b200 kernels are built for hidden_features={HIDDEN_FEATURES}, got {self.siren_hidden_features}")
        if not 2 <= self.siren_hidden_layers <= 7:
            raise NotImplementedError("reni_b200 FiLM kernels support 2..7 FiLM layers")
        if not 1 <= self.out_features <= 3:
            raise NotImplementedError("reni_b200 kernels support out_features <= 3")

    def c_config(self) -> RENIConfig:
        """The core sees L = siren_hidden_layers - 1 modulated 256x256 layers with omega = 1 (the per-map frequencies
        take the place of omega_0) and a linear output layer; exp is applied by the caller."""
        self.validate()
        return RENIConfig(self.ndims, _lib.EQUIVARIANCE[self.equivariance], self.siren_hidden_features,
                          self.siren_hidden_layers - 1, self.out_features, 1,
                          1 if self.output_activation == "tanh" else 0, 1.0, 1.0)


def _film_permap_flag(P: int) -> int:
    """FLAG_FILM_PERMAP when the FiLM core can run on per-map weight images (include/reni_b200.h,
    reni_film_prepare_maps): every unit of four 128-direction tiles must lie inside one map.  RENI_FILM_PERMAP=0 keeps
    the in-epilogue modulation (A/B switch)."""
    if os.environ.get("RENI_FILM_PERMAP", "1") == "0":
        return 0
    return _lib.FLAG_FILM_PERMAP if P % 512 == 0 else 0


def _film_prepare_maps(lib, cfg, filmc, weights, biases, B: int, P: int, ws: "Workspace", flags: int, dev) -> None:
    if flags & _lib.FLAG_FILM_PERMAP:
        rc = _call(dev, lib.reni_film_prepare_maps, C.byref(cfg), _vp(filmc), _ptr_array([weights[0]] + weights),
                                        _ptr_array([biases[0]] + biases), B, P, _vp(ws.view), ws.nbytes, flags,
                                        _stream(dev))
        _lib.check(rc, "reni_film_prepare_maps")


class _FilmCoreFunction(torch.autograd.Function):
    """out = core(mc, film, D; W_1..W_L, b_1..b_L, W_out, b_out) with gradients for mc, film and the parameters."""

    @staticmethod
    def forward(ctx, spec: FilmSpec, inference_ws: Workspace, mc, film, D, *params):
        lib = _lib.load()
        cfg = spec.c_config()
        L = spec.siren_hidden_layers - 1
        weights = [_f32c(p) for p in params[0::2]]   # W_1 .. W_L, W_out
        biases = [_f32c(p) for p in params[1::2]]
        assert len(weights) == L + 1 and len(biases) == L + 1
        dev = _require_cuda(mc, film, D, *weights, *biases)
        mcc, filmc = _f32c(mc), _f32c(film)
        B = mcc.shape[0]
        if tuple(mcc.shape) != (B, 5, HIDDEN_FEATURES) or tuple(filmc.shape) != (B, L, 2, HIDDEN_FEATURES):
            raise ValueError(f"mc must be (B,5,256) and film (B,{L},2,256); got {tuple(mcc.shape)}, {tuple(filmc.shape)}")
        Dc, d_bs = _batch_stride(D, B, "directions")
        P = Dc.shape[1]
        need_in = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        need_dw = any(ctx.needs_input_grad[5:])
        flags = _lib.FLAG_FILM | _film_permap_flag(P)
        if need_in or need_dw:
            flags |= FLAG_SAVE_FOR_BACKWARD | (FLAG_NEED_DW if need_dw else 0)
        ws = Workspace() if (flags & FLAG_SAVE_FOR_BACKWARD) else inference_ws
        ws.ensure(workspace_bytes(cfg, B, P, flags), dev)
        # weight images: slot 0 of the parameter arrays (the first layer) is not used by the core
        prepare_weights(cfg, [weights[0]] + weights, [biases[0]] + biases, ws, dev)
        _film_prepare_maps(lib, cfg, filmc, weights, biases, B, P, ws, flags, dev)
        out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
        rc = _call(dev, lib.reni_film_forward, C.byref(cfg), _vp(mcc), _vp(filmc), _vp(Dc), d_bs, B, P, _vp(out), _vp(ws.view),
                                   ws.nbytes, flags, _stream(dev))
        _lib.check(rc, "reni_film_forward")
        if flags & FLAG_SAVE_FOR_BACKWARD:
            ctx.spec, ctx.ws, ctx.flags, ctx.d_bs = spec, ws, flags, d_bs
            ctx.save_for_backward(filmc, Dc, out, *weights, *biases)
        if spec.out_features != 3:
            return out[:, :, : spec.out_features]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        spec: FilmSpec = ctx.spec
        cfg = spec.c_config()
        L = spec.siren_hidden_layers - 1
        saved = ctx.saved_tensors
        filmc, Dc, out = saved[0], saved[1], saved[2]
        weights, biases = list(saved[3:3 + L + 1]), list(saved[3 + L + 1:3 + 2 * (L + 1)])
        dev = filmc.device
        B, P = filmc.shape[0], Dc.shape[1]
        g = _f32c(grad_out)
        if g.shape[2] != 3:
            gp = torch.zeros(B, P, 3, device=dev)
            gp[:, :, : g.shape[2]] = g
            g = gp
        need_dw = bool(ctx.flags & FLAG_NEED_DW)
        d_mc = torch.empty(B, 5, HIDDEN_FEATURES, device=dev, dtype=torch.float32)
        d_film = torch.empty_like(filmc)
        dW = [torch.zeros_like(w) for w in weights] if need_dw else None
        db = [torch.zeros_like(b) for b in biases] if need_dw else None
        rc = _call(dev, lib.reni_film_backward, 
            C.byref(cfg), _vp(filmc), _vp(Dc), ctx.d_bs, _ptr_array([weights[0]] + weights),
            _ptr_array([biases[0]] + biases), B, P, _vp(out), _vp(g), _vp(d_mc), _vp(d_film),
            _ptr_array([None] + dW) if need_dw else None, _ptr_array([None] + db) if need_dw else None,
            _vp(ctx.ws.view), ctx.ws.nbytes, ctx.flags, _stream(dev))
        _lib.check(rc, "reni_film_backward")
        ctx.ws = None  # release the stash
        grads: List[Optional[torch.Tensor]] = []
        for i in range(L + 1):
            grads.append(dW[i] if need_dw and ctx.needs_input_grad[5 + 2 * i] else None)
            grads.append(db[i] if need_dw and ctx.needs_input_grad[6 + 2 * i] else None)
        return (None, None, d_mc if ctx.needs_input_grad[2] else None, d_film if ctx.needs_input_grad[3] else None,
                None, *grads)


def film_decode_core(spec: FilmSpec, inference_ws: Workspace, mc: torch.Tensor, film: torch.Tensor, D: torch.Tensor,
                     params: Sequence[torch.Tensor]) -> torch.Tensor:
    """``params`` = [W_1, b_1, ..., W_L, b_L, W_out, b_out] (net.1.. and final_layer of the reference's FiLM module).
    Differentiable w.r.t. mc, film and the parameters; the output activation (tanh in-kernel) excludes ``exp``."""
    spec.validate()
    if torch.is_grad_enabled() and (mc.requires_grad or film.requires_grad or any(p.requires_grad for p in params)):
        return _FilmCoreFunction.apply(spec, inference_ws, mc, film, D, *params)
    with torch.no_grad():
        return _FilmCoreFunction.forward(_NoGradCtx(len(params) + 1), spec, inference_ws, mc, film, D, *params)


@dataclass
class FilmStepResult:
    """Fused FiLM step around the core: loss parts, radiance, and the gradients the per-map stage continues from."""

    loss: torch.Tensor        # mse + beta * cosine  (the prior / KLD terms are per-map terms of the caller)
    mse_loss: torch.Tensor
    cosine_loss: torch.Tensor
    out: torch.Tensor         # (B, P, 3)
    d_mc: torch.Tensor        # (B, 5, 256)
    d_film: torch.Tensor      # (B, L, 2, 256)
    dW: Optional[List[torch.Tensor]]   # [W_1 .. W_L, W_out]
    db: Optional[List[torch.Tensor]]


def film_loss_forward_backward(spec: FilmSpec, ws: Workspace, mc: torch.Tensor, film: torch.Tensor, D: torch.Tensor,
                               target: torch.Tensor, sineweight: torch.Tensor, params: Sequence[torch.Tensor],
                               beta: float = 0.0, use_cosine: bool = False, need_dw: bool = True,
                               grad_weights: Optional[Sequence[torch.Tensor]] = None,
                               grad_biases: Optional[Sequence[torch.Tensor]] = None) -> FilmStepResult:
    """Forward + WeightedMSE (+ beta * WeightedCosineSimilarity) + backward of the FiLM core in one library call
    (loss_functions.py:6-32; RENI_module.py:105-134 without the per-map terms).  ``params`` as in
    ``film_decode_core``; ``grad_weights`` / ``grad_biases`` (one per core parameter pair) are ACCUMULATED into."""
    spec.validate()
    if spec.output_activation == "exp":
        raise NotImplementedError("the fused FiLM step covers output_activation None / 'tanh'; use the autograd path")
    lib = _lib.load()
    cfg = spec.c_config()
    L = spec.siren_hidden_layers - 1
    weights = [_f32c(p) for p in params[0::2]]
    biases = [_f32c(p) for p in params[1::2]]
    dev = _require_cuda(mc, film, D, target, sineweight, *weights, *biases)
    mcc, filmc = _f32c(mc), _f32c(film)
    B = mcc.shape[0]
    if tuple(mcc.shape) != (B, 5, HIDDEN_FEATURES) or tuple(filmc.shape) != (B, L, 2, HIDDEN_FEATURES):
        raise ValueError(f"mc must be (B,5,256) and film (B,{L},2,256); got {tuple(mcc.shape)}, {tuple(filmc.shape)}")
    Dc, d_bs = _batch_stride(D, B, "directions")
    swc, sw_bs = _batch_stride(sineweight, B, "sineweight")
    P = Dc.shape[1]
    tc = _f32c(target)
    if tuple(tc.shape) != (B, P, 3):
        raise ValueError(f"target must have shape {(B, P, 3)}, got {tuple(tc.shape)}")
    flags = _lib.FLAG_FILM | FLAG_SAVE_FOR_BACKWARD | FLAG_LOSS | (FLAG_NEED_DW if need_dw else 0) | _film_permap_flag(P)
    ws.ensure(workspace_bytes(cfg, B, P, flags), dev)
    prepare_weights(cfg, [weights[0]] + weights, [biases[0]] + biases, ws, dev)
    _film_prepare_maps(lib, cfg, filmc, weights, biases, B, P, ws, flags, dev)
    out = torch.empty(B, P, 3, device=dev, dtype=torch.float32)
    loss = torch.empty(4, device=dev, dtype=torch.float32)
    d_mc = torch.empty(B, 5, HIDDEN_FEATURES, device=dev, dtype=torch.float32)
    d_film = torch.empty_like(filmc)
    dW = db = None
    if need_dw:
        dW = list(grad_weights) if grad_weights is not None else [torch.zeros_like(w) for w in weights]
        db = list(grad_biases) if grad_biases is not None else [torch.zeros_like(b) for b in biases]
    rc = _call(dev, lib.reni_film_loss_forward_backward, 
        C.byref(cfg), _vp(mcc), _vp(filmc), _vp(Dc), d_bs, _ptr_array([weights[0]] + weights),
        _ptr_array([biases[0]] + biases), B, P, _vp(tc), _vp(swc), sw_bs, float(beta), 1 if use_cosine else 0, _vp(out),
        _vp(loss), _vp(d_mc), _vp(d_film), _ptr_array([None] + dW) if need_dw else None,
        _ptr_array([None] + db) if need_dw else None, _vp(ws.view), ws.nbytes, flags, _stream(dev))
    _lib.check(rc, "reni_film_loss_forward_backward")
    return FilmStepResult(loss[0], loss[1], loss[3], out, d_mc, d_film, dW, db)
