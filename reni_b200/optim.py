"""Fused Adam for the RENI training / latent-fit loops.

The reference builds ``torch.optim.Adam(params, lr)`` (src/lightning/RENI_module.py:185-192: the configured betas are
never passed, so betas = (0.9, 0.999), eps = 1e-8, no weight decay) and steps it densely over every parameter,
including the whole latent table.  ``FusedAdam`` does the same arithmetic in ONE launch of ``reni_adam_step`` over all
parameter segments; it is a ``torch.optim.Optimizer``, so the reference's per-epoch ``ExponentialLR``
(RENI_module.py:212-214) drives ``param_groups[i]["lr"]`` unchanged.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise RuntimeError("reni_b200.FusedAdam runs on CUDA only (no CPU fallback)")
            segs = (_lib.AdamSegment * len(ps))()
            keep = []
            for i, p in enumerate(ps):
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam expects contiguous fp32 parameters")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                segs[i] = _lib.AdamSegment(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                           st["exp_avg_sq"].data_ptr(), p.numel())
            step = group.get("_step")
            if step is None:  # one device-side step counter per group (all its parameters step together)
                step = group["_step"] = torch.zeros(1, dtype=torch.int32, device=dev)
            rc = lib.reni_adam_step(segs, len(ps), C.c_void_p(step.data_ptr()), float(group["lr"]),
                                    float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc, "reni_adam_step")
            # the kernel wrote through raw pointers: tell autograd (and the decoder's weight-image cache, which keys on
            # the version counters) that the parameters changed in place
            torch.autograd.graph.increment_version(ps)
        return loss
