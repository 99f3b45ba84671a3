"""Training / latent-optimisation step of the reference, on the fused kernels.

Restates what one Lightning iteration does around the decoder (reference:
src/lightning/RENI_module.py:80-146 ``training_step``, :168-195 optimiser, run.py:97-108
DDPStrategy; examples.ipynb cell 4 is the same loop by hand):

    imgs (B,3,H,W) -> (B,P,3);  Z = model.Z[idx] (or mu[idx] / sample_latent(idx));
    out = model(Z, directions);  loss = criterion(out, imgs, sineweight [* mask], ...);
    loss.backward();  [DDP: all-reduce(avg) of every gradient];  Adam.step()

Here forward + loss + backward is ONE library call (``loss_forward_backward``); the decoder
weight gradients land in a single flat fp32 buffer that is all-reduced with ONE NCCL call;
latent gradients stay on the rank that owns the maps.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import functional as F_
from .functional import Workspace
from .geometry import get_directions, get_sineweight
from .film import RENIVariationalAutoDecoderFiLM, _FilmDecoderBase
from .losses import KLD, RENITestLoss, RENITrainLoss, RENIVADTrainLoss
from .models import RENIAutoDecoder, RENIVariationalAutoDecoder, _DecoderBase
from .optim import FusedAdam


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of maps owned by ``rank`` (maps are independent units; no data-path collective)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradBuffer:
    """One contiguous fp32 buffer holding every decoder-weight gradient (680 707 floats at N=36), with
    per-parameter views used both as the kernels' accumulation targets and as ``param.grad``.

    Exchange (DDP semantics, run.py:97: average over ranks of the per-rank batch-summed gradients), chosen once:
      * ``"multicast"`` / ``"p2p"``: the buffer lives in symmetric memory (torch.distributed._symmetric_memory does the
        peer mapping) and ``reni_allreduce`` reduces it in place through NVLink -- multimem.ld_reduce / multimem.st on
        the NVSwitch multicast address, or a two-shot exchange over the peer pointers.  A plain kernel on the
        caller's stream: it can be captured in the step graph right behind the gradient kernels (``capturable``).
      * ``"nccl"``: one ``ncclAllReduce`` (gloo: sum + divide) -- the fallback and the CPU-test path.
    ``RENI_EXCHANGE=nccl|p2p|multicast`` forces a choice."""

    def __init__(self, params: Sequence[torch.Tensor], group=None, symmetric: Optional[bool] = None):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.group = group
        self.exchange = "none"
        self.capturable = True  # (nothing to exchange at world size 1)
        self._symm = None
        self.flat = None
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            self.exchange, self.capturable = "nccl", False
            want = os.environ.get("RENI_EXCHANGE", "")
            if symmetric is not False and want != "nccl" and dev.type == "cuda" and dist.get_backend(group) == "nccl":
                try:
                    self._setup_symmetric(n, dev, group, world, want)
                except Exception as e:  # symmetric memory unavailable (driver, topology, torch build): NCCL does it
                    self._symm = None
                    self.flat = None
                    self.symmetric_error = repr(e)
                    if symmetric is True or want in ("p2p", "multicast"):
                        raise
        if self.flat is None:
            self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def _setup_symmetric(self, n: int, dev, group, world: int, want: str) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        lib = _lib.load()
        pg = group if group is not None else dist.group.WORLD
        n4 = (n + 3) // 4 * 4
        buf = symm_mem.empty(n4, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, pg)
        nflag = int(lib.reni_allreduce_flag_bytes()) // 4
        flags = symm_mem.empty(nflag, dtype=torch.int32, device=dev)
        flags.zero_()
        fh = symm_mem.rendezvous(flags, pg)
        if getattr(hdl, "offset", 0) or getattr(fh, "offset", 0):
            raise RuntimeError("symmetric allocation at a non-zero offset")
        buf.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's flag block is zero before anybody signals
        try:
            mc = int(hdl.multicast_ptr) if want != "p2p" else 0
        except Exception:
            mc = 0
        if want == "multicast" and not mc:
            raise RuntimeError("no multicast address for the symmetric gradient buffer")
        self._symm = dict(buf=buf, hdl=hdl, flags=flags, fh=fh, mc=mc, n4=n4, world=world, rank=dist.get_rank(group),
                          epoch=torch.zeros(nflag, dtype=torch.int32, device=dev),
                          status=torch.zeros(1, dtype=torch.int32, device=dev))
        self.flat = buf[:n]
        self.exchange = "multicast" if mc else "p2p"
        self.capturable = True

    def zero_(self) -> None:
        self.flat.zero_()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce_mean(self, group=None) -> None:
        """DDP semantics (run.py:97): average over ranks of the per-rank batch-summed gradients."""
        if self.exchange == "none" or not (dist.is_available() and dist.is_initialized()):
            return
        group = group if group is not None else self.group
        if self._symm is not None:
            import ctypes as C

            from . import _lib

            sm = self._symm
            dev = self.flat.device
            with torch.cuda.device(dev):
                rc = _lib.load().reni_allreduce(
                    C.c_void_p(int(sm["hdl"].buffer_ptrs_dev)), C.c_void_p(int(sm["fh"].buffer_ptrs_dev)),
                    C.c_void_p(sm["mc"]), sm["n4"], sm["rank"], sm["world"], 1.0 / sm["world"],
                    C.c_void_p(sm["epoch"].data_ptr()), C.c_void_p(sm["status"].data_ptr()),
                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(rc, "reni_allreduce")
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)

    def exchange_failed(self) -> bool:
        """True if a rank gave up waiting for a peer inside reni_allreduce (host sync; for tests and diagnostics)."""
        return self._symm is not None and bool(int(self._symm["status"].item()))


class RENITrainer:
    """One task of the reference's ``RENI`` LightningModule, minus Lightning.

    task: "FIT_DECODER" (all parameters, RENITrainLoss / RENIVADTrainLoss) or "FIT_LATENT"
    (frozen decoder, RENITestLoss with optional mask; RENI_module.py:295-336)."""

    def __init__(self, model: _DecoderBase, task: str, sidelen: int, lr: float = 1e-5,
                 prior_loss_weight: float = 1e-7, cosine_similarity_weight: float = 1e-4,
                 kld_weighting: float = 1e-4, mask: Optional[torch.Tensor] = None, process_group=None,
                 ddp_latent_scaling: bool = True, cuda_graph: bool = False, latent_sync: str = "local",
                 analytic_grid: bool = False):
        if task not in ("FIT_DECODER", "FIT_LATENT"):
            raise NotImplementedError("FIT_INVERSE needs the PyTorch3D renderer and is out of scope for this path")
        self.model = model
        self.task = task
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.ddp_latent_scaling = ddp_latent_scaling
        # Latent tables under data parallelism.  "local" (default): maps are sharded, a rank only ever updates the rows
        # of the maps it owns (shard_range) and the tables are NOT kept identical across ranks -- feed each rank its
        # own shard and use gather_latents() for a checkpoint.  "replicated": the reference's DDP behaviour, the dense
        # table gradient is all-reduced too, every rank holds the same table and any sampler works.
        if latent_sync not in ("local", "replicated"):
            raise ValueError("latent_sync must be 'local' or 'replicated'")
        self.latent_sync = latent_sync
        dev = next(model.parameters()).device
        self.device = dev
        # analytic_grid=True: the kernels compute get_directions / get_sineweight from the pixel index and take the mask
        # as one bit per pixel, instead of reading the (1, P, 3) arrays and a torch `sineweight * mask` per step
        # (Cond-by-Concat decoders; utils.py:46-78, RENI_module.py:89-94)
        self.analytic_grid = bool(analytic_grid)
        self.set_resolution(sidelen)
        self.mask = mask.to(dev) if mask is not None else None
        self.alpha = prior_loss_weight
        self.beta = cosine_similarity_weight
        self.kld_weighting = kld_weighting
        self.is_vad = isinstance(model, (RENIVariationalAutoDecoder, RENIVariationalAutoDecoderFiLM))
        # FiLM decoders (the reference's default conditioning) step through autograd: the per-map stage (mapping
        # network, hoisted first layer) is torch, the per-direction stage is reni_film_forward / reni_film_backward
        self.is_film = isinstance(model, _FilmDecoderBase)
        self.fixed = task == "FIT_LATENT"
        self._ws = Workspace()
        # cuda_graph=True: the whole step (weight images, prologue, fwd, loss, bwd, grad exchange) is captured once per
        # batch shape and replayed from static input buffers -- ~12 launches become one, which matters for a <1 ms step
        self.cuda_graph = bool(cuda_graph)
        self._graphs: Dict[tuple, tuple] = {}
        # prefetch(): double-buffered device staging filled on a copy stream, so the host->device copy of batch i+1
        # runs under the step of batch i (what a pinned-memory DataLoader with prefetching gives the reference)
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._stage: Dict[tuple, list] = {}
        self._stage_next = 0
        self._pending: Optional[tuple] = None
        # optimiser: Adam(lr) with default betas -- the cfg betas are never passed (RENI_module.py:191-192)
        if self.fixed:
            opt_params = [model.mu] if self.is_vad else [model.Z]  # RENI_module.py:178-183
            self.flat = None
        else:
            opt_params = list(model.parameters())
            latents = {id(p) for p in (getattr(model, n, None) for n in ("Z", "mu", "log_var")) if p is not None}
            self.flat = FlatGradBuffer([p for p in model.parameters() if id(p) not in latents] if self.is_film
                                       else model.decoder_parameters(), group=process_group)
        self.optimizer = FusedAdam(opt_params, lr=lr)  # one launch; same arithmetic as torch.optim.Adam(params, lr)

    # -- multi-resolution curriculum hook (callbacks.py:11-29 doubles the resolution at curriculum epochs)
    def set_resolution(self, sidelen: int) -> None:
        self.sidelen = sidelen
        self.directions = get_directions(sidelen).to(self.device)   # (1, P, 3), shared by every map
        self.sineweight = get_sineweight(sidelen).to(self.device)   # (1, P, 3)

    def on_train_epoch_end(self, current_epoch: int, curriculum: Sequence[int], mask_fn=None) -> bool:
        """MultiResTrainingCallback.on_train_epoch_end (callbacks.py:11-25): when ``current_epoch + 1`` is a curriculum
        epoch, double the resolution of the directions, the sine weights and (through ``mask_fn(sidelen)``, the
        reference re-reads the PNG with get_mask) the mask.  Returns True when the resolution changed, i.e. when the
        caller must also switch its dataset to the doubled resolution (``dataset.double_resolution()``).  Captured
        CUDA graphs are keyed by batch shape, so the new resolution simply captures its own graph."""
        if current_epoch + 1 not in curriculum:
            return False
        self.set_resolution(2 * self.sidelen)
        if self.mask is not None:
            if mask_fn is None:
                raise ValueError("a masked task needs mask_fn(sidelen) to rebuild the mask at the new resolution")
            self.mask = mask_fn(self.sidelen).to(self.device)
        return True

    def exponential_lr(self, lr_start: float, lr_end: float, epochs: int):
        """Per-epoch ExponentialLR with gamma = exp(ln(lr_end/lr_start)/epochs) (RENI_module.py:212-214)."""
        gamma = math.exp(math.log(lr_end / lr_start) / epochs)
        return torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=gamma)

    def _latent_table(self) -> torch.nn.Parameter:
        return self.model.mu if self.is_vad else self.model.Z

    def training_step(self, batch, batch_idx: int = 0) -> Dict[str, torch.Tensor]:
        """Same inputs and returned keys as RENI_module.training_step; gradients are left in ``.grad``."""
        if self.is_film:
            return self._film_step(batch)
        if self.cuda_graph:  # (the VAD sampler's noise is a torch.randn inside the capture: graph-safe Philox offsets)
            return self._graphed_step(batch)
        return self._eager_step(batch)

    def _film_step(self, batch) -> Dict[str, torch.Tensor]:
        """RENI_module.training_step (:80-146) for a FiLM decoder, op for op: model(Z, D) -> criterion -> backward.
        Decoder-side gradients accumulate straight into the flat all-reduce buffer (its views are the ``.grad``s).
        With ``cuda_graph=True`` the whole autograd step (mapping network, fused core, loss, backward: ~150 launches
        for ~1 ms of GPU work) is captured once per batch shape and replayed from static inputs."""
        slot = self._take_prefetched(batch)
        imgs, idx = batch
        idx = torch.as_tensor(idx, dtype=torch.long)
        graphed = self.cuda_graph and not (self.is_vad and not self.fixed)  # (the VAD sampler draws from torch's RNG)
        if not graphed:
            if slot is not None:
                imgs, idx = slot["imgs"].clone(), slot["idx"].clone()
                slot["free"].record(torch.cuda.current_stream(self.device))
            log = self._film_body(imgs.to(self.device, non_blocking=True), idx.to(self.device, non_blocking=True))
        else:
            key = ("film", tuple(imgs.shape), imgs.dtype, int(idx.numel()))
            entry = self._graphs.get(key)
            if entry is None:
                s_imgs = torch.empty(imgs.shape, dtype=imgs.dtype, device=self.device)
                s_idx = torch.empty(idx.shape, dtype=torch.long, device=self.device)
                s_imgs.copy_(imgs)
                s_idx.copy_(idx)
                # a captured graph holds raw pointers into its workspace: each graph owns one (a shared, growing
                # workspace would be freed under it by a later, larger batch shape)
                gws = Workspace()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):  # warm-up outside capture (lazy initialisation, cuBLAS workspaces)
                    for _ in range(3):
                        gws.prepared_key = None
                        self._film_body(s_imgs, s_idx, ws=gws)
                torch.cuda.current_stream(self.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                gws.prepared_key = None  # the fp32 -> fp16 weight images are rebuilt inside every replayed step
                with torch.cuda.graph(graph):
                    log = self._film_body(s_imgs, s_idx, ws=gws)
                grads = [(p, p.grad) for p in self.model.parameters()]
                entry = (graph, s_imgs, s_idx, log, grads, self.last_output, gws)
                self._graphs[key] = entry
            graph, s_imgs, s_idx, log, grads, last_out, _gws = entry
            src_imgs, src_idx = (slot["imgs"], slot["idx"]) if slot is not None else (imgs, idx)
            s_imgs.copy_(src_imgs, non_blocking=True)
            s_idx.copy_(src_idx, non_blocking=True)
            if slot is not None:
                slot["free"].record(torch.cuda.current_stream(self.device))
            graph.replay()
            for p, g in grads:  # the replay refreshed the captured gradient tensors in place
                p.grad = g
            self.last_output = last_out
        if self.world_size > 1 and self.ddp_latent_scaling:
            for p in (getattr(self.model, n, None) for n in ("Z", "mu", "log_var")):
                if p is not None and p.grad is not None:
                    p.grad.div_(self.world_size)
        if self.flat is not None:
            self.flat.all_reduce_mean(self.group)
        self._sync_latent_grads()
        return log

    def _film_body(self, imgs: torch.Tensor, idx: torch.Tensor, ws: Optional[Workspace] = None) -> Dict[str, torch.Tensor]:
        ws = ws if ws is not None else self._ws
        B = imgs.shape[0]
        imgs = imgs.permute(0, 2, 3, 1).reshape(B, -1, 3)
        sw = self.sineweight if self.mask is None else self.sineweight * self.mask
        sw = sw.expand(B, -1, -1)
        D = self.directions.expand(B, -1, -1)
        model = self.model
        for p in model.parameters():
            p.grad = None
        if self.flat is not None:
            self.flat.zero_()
            self.flat.attach()
        if self.is_vad and not self.fixed:
            Z, mu, log_var = model.sample_latent(idx)
        else:
            # (== table[idx]; index_select's backward is one index_add_ instead of advanced indexing's sort + scatter)
            Z = self._latent_table().index_select(0, idx)
        if model.output_activation != "exp":
            # fused core step: forward + WeightedMSE (+ cosine) + backward in one library call; the per-map stage
            # (mapping network, hoisted first layer, prior / KLD) is differentiated by autograd from d_mc / d_film
            need_dw = not self.fixed
            sinks = None
            if need_dw:  # the native per-map stage adds its parameter gradients straight into the flat buffer's views
                by_id0 = {id(p): v for p, v in zip(self.flat.params, self.flat.views)}
                sinks = [by_id0[id(p)] for p in model._map_params()]
            mc, film = model.map_level(Z.float(), grad_sinks=sinks)
            core = model.core_parameters()
            views = None
            if need_dw:
                by_id = {id(p): v for p, v in zip(self.flat.params, self.flat.views)}
                views = [by_id[id(p)] for p in core]
            fit_latent = self.task == "FIT_LATENT"
            res = F_.film_loss_forward_backward(
                model.spec, ws, mc.detach(), film.detach(), self.directions, imgs,
                self.sineweight if self.mask is None else self.sineweight * self.mask, core,
                beta=self.beta if fit_latent else 0.0, use_cosine=fit_latent, need_dw=need_dw,
                grad_weights=views[0::2] if need_dw else None, grad_biases=views[1::2] if need_dw else None)
            roots, grads = [mc, film], [res.d_mc, res.d_film]
            if fit_latent:
                prior = self.alpha * torch.sum(Z ** 2)                      # RENITestLoss (loss_functions.py:68)
                roots.append(prior)
                grads.append(torch.ones_like(prior))
                log = {"loss": res.loss + prior.detach(), "mse_loss": res.mse_loss, "prior_loss": prior.detach(),
                       "cosine_loss": res.cosine_loss}
            elif self.is_vad:
                kld = self.kld_weighting * KLD(mu, log_var, Z_dims=model.ndims * 3)   # RENI_module.py:312-315
                roots.append(kld)
                grads.append(torch.ones_like(kld))
                log = {"loss": res.loss + kld.detach(), "mse_loss": res.mse_loss, "kld_loss": kld.detach()}
            else:
                log = {"loss": res.loss}
            torch.autograd.backward(roots, grads)
            self.last_output = res.out
            return log
        out = model(Z, D)
        if self.task == "FIT_LATENT":
            loss, mse, prior, cos = RENITestLoss(alpha=self.alpha, beta=self.beta)(out, imgs, sw, Z)
            log = {"loss": loss.detach(), "mse_loss": mse.detach(), "prior_loss": prior.detach(),
                   "cosine_loss": cos.detach()}
        elif self.is_vad:
            loss, mse, kld = RENIVADTrainLoss(beta=self.kld_weighting, Z_dims=model.ndims * 3)(out, imgs, sw, mu, log_var)
            log = {"loss": loss.detach(), "mse_loss": mse.detach(), "kld_loss": kld.detach()}
        else:
            loss = RENITrainLoss()(out, imgs, sw)
            log = {"loss": loss.detach()}
        loss.backward()
        self.last_output = out.detach()
        return log

    def prefetch(self, batch) -> None:
        """Start copying ``batch`` (pinned host tensors) to the device on a side stream.  The next
        ``training_step(batch)`` called with the same tensors consumes the staged copy instead of copying itself."""
        imgs, idx = batch
        idx = torch.as_tensor(idx, dtype=torch.long)
        key = (tuple(imgs.shape), imgs.dtype, int(idx.numel()))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        slots = self._stage.get(key)
        if slots is None:
            slots = []
            for _ in range(2):
                slots.append({"imgs": torch.empty(imgs.shape, dtype=imgs.dtype, device=self.device),
                              "idx": torch.empty(idx.shape, dtype=torch.long, device=self.device),
                              "ready": torch.cuda.Event(), "free": torch.cuda.Event()})
                slots[-1]["free"].record(torch.cuda.current_stream(self.device))
            self._stage[key] = slots
        slot = slots[self._stage_next]
        self._stage_next ^= 1
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(slot["free"])  # the step that last read this slot has consumed it
            slot["imgs"].copy_(imgs, non_blocking=True)
            slot["idx"].copy_(idx, non_blocking=True)
            slot["ready"].record(self._copy_stream)
        self._pending = (batch[0], batch[1], slot)

    def _take_prefetched(self, batch):
        """Device tensors of ``batch`` if it is the pending prefetch (waits for the copy on the current stream)."""
        pend = self._pending
        if pend is None or pend[0] is not batch[0] or pend[1] is not batch[1]:
            return None
        self._pending = None
        slot = pend[2]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return slot

    def _graphed_step(self, batch) -> Dict[str, torch.Tensor]:
        slot = self._take_prefetched(batch)
        imgs, idx = batch
        idx = torch.as_tensor(idx, dtype=torch.long)
        self._check_shard(idx)
        key = (tuple(imgs.shape), imgs.dtype, int(idx.numel()))
        entry = self._graphs.get(key)
        # the gradient exchange is part of the graph when it is a plain kernel (reni_allreduce over symmetric memory);
        # an NCCL all-reduce is launched behind each replay instead
        in_graph = self.flat is None or self.flat.capturable
        if entry is None:
            s_imgs = torch.empty(imgs.shape, dtype=imgs.dtype, device=self.device)
            s_idx = torch.empty(idx.shape, dtype=torch.long, device=self.device)
            s_imgs.copy_(imgs)
            s_idx.copy_(idx)
            # a captured graph holds raw pointers into its workspace: each graph owns one (a shared, growing workspace
            # would be freed under it by a later, larger batch shape)
            gws = Workspace()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside capture: lazy allocations, function attributes
                for _ in range(2):
                    gws.prepared_key = None
                    self._eager_step((s_imgs, s_idx), exchange=in_graph, ws=gws)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            gws.prepared_key = None  # the fp32 -> fp16 weight conversion is part of every replayed step
            with torch.cuda.graph(graph):
                log = self._eager_step((s_imgs, s_idx), exchange=in_graph, ws=gws)
            grads = [(p, p.grad) for p in (getattr(self.model, n, None) for n in ("Z", "mu", "log_var"))
                     if p is not None and p.grad is not None]
            entry = (graph, s_imgs, s_idx, log, grads, gws)
            self._graphs[key] = entry
        graph, s_imgs, s_idx, log, grads, _gws = entry
        if slot is not None:  # staged by prefetch(): device -> device into the graph's static inputs
            s_imgs.copy_(slot["imgs"], non_blocking=True)
            s_idx.copy_(slot["idx"], non_blocking=True)
            slot["free"].record(torch.cuda.current_stream(self.device))
        else:
            s_imgs.copy_(imgs, non_blocking=True)
            s_idx.copy_(idx, non_blocking=True)
        graph.replay()
        # the replay refreshed the captured gradient tensors in place; make sure the parameters still point at them
        for p, g in grads:
            p.grad = g
        if self.flat is not None:
            if not in_graph:
                self.flat.all_reduce_mean(self.group)  # the ONE exchange step of data-parallel training
            self.flat.attach()
        return log

    def _eager_step(self, batch, exchange: bool = True, ws: Optional[Workspace] = None) -> Dict[str, torch.Tensor]:
        ws = ws if ws is not None else self._ws
        slot = self._take_prefetched(batch)
        imgs, idx = batch
        if slot is not None:
            imgs, idx = slot["imgs"].clone(), slot["idx"].clone()
            slot["free"].record(torch.cuda.current_stream(self.device))
        else:
            self._check_shard(idx)
        imgs = imgs.to(self.device, non_blocking=True)
        B = imgs.shape[0]
        imgs = imgs.permute(0, 2, 3, 1).reshape(B, -1, 3)  # (B,C,H,W) -> (B,P,3)   RENI_module.py:83-84
        sw = self.sineweight if self.mask is None else self.sineweight * self.mask  # :90-94
        idx = torch.as_tensor(idx, device=self.device, dtype=torch.long)
        model = self.model
        table = self._latent_table()
        sampled = self.is_vad and not self.fixed
        latent_scale = 1.0 / self.world_size if (self.world_size > 1 and self.ddp_latent_scaling) else 1.0
        if sampled:
            # RENI_module.py:100-101 -> sample_latent (RENI.py:329-335): the noise is torch's (same generator use as
            # randn_like(std)), the sample itself one library call; nothing here needs an autograd graph
            eps = torch.randn(B, model.ndims, 3, device=self.device, dtype=torch.float32)
            Zin = torch.empty_like(eps)
            self._vad_call("reni_vad_sample", model.mu, model.log_var, idx, eps, B, model.ndims * 3, Zin)
        else:
            Zin = table.detach()[idx]  # :98,:103
        need_dw = not self.fixed
        if need_dw:
            self.flat.zero_()
        if self.task == "FIT_LATENT":
            alpha, beta, use_cos = self.alpha, self.beta, True      # RENITestLoss  (:126-128)
        else:
            alpha, beta, use_cos = 0.0, 0.0, False                  # RENITrainLoss (:117)
        grid_kw = {}
        D_arg, sw_arg = self.directions, sw
        if self.analytic_grid:
            from .geometry import pack_mask_bits

            D_arg, sw_arg = None, None
            if self.mask is not None:
                if getattr(self, "_mask_bits_for", None) is not self.mask:
                    self._mask_bits, self._mask_bits_for = pack_mask_bits(self.mask), self.mask
                grid_kw["mask_bits"] = self._mask_bits
        res = F_.loss_forward_backward(
            model.spec, ws, Zin, D_arg, imgs, sw_arg, model.decoder_weights(), model.decoder_biases(),
            alpha=alpha, beta=beta, use_cosine=use_cos, need_dw=need_dw, **grid_kw,
            grad_weights=self.flat.views[0::2] if need_dw else None,
            grad_biases=self.flat.views[1::2] if need_dw else None)
        log: Dict[str, torch.Tensor] = {"loss": res.loss}
        if sampled:
            # RENIVADTrainLoss (loss_functions.py:47-58, beta = KLD_WEIGHTING, Z_dims = 3N: RENI_module.py:312-315): the KLD
            # value and BOTH terms' gradients for mu / log_var in one launch; the reference's DDP averages the whole
            # gradient of the replicated tables, so the KLD part is divided by the world size like the MSE part
            model.mu.grad = torch.zeros_like(model.mu)
            model.log_var.grad = torch.zeros_like(model.log_var)
            kld = torch.zeros((), device=self.device, dtype=torch.float32)
            nz = model.ndims * 3
            self._vad_call("reni_vad_backward", model.mu, model.log_var, idx, eps, res.dZ, B, nz,
                           float(self.kld_weighting) / nz, float(latent_scale), model.mu.grad, model.log_var.grad, kld)
            log = {"loss": res.loss + kld, "mse_loss": res.mse_loss, "kld_loss": kld}
        else:
            # the reference replicates the latent table and DDP averages its (mostly zero) gradient over ranks
            dZ = res.dZ * latent_scale if latent_scale != 1.0 else res.dZ
            g = torch.zeros_like(table)
            g.index_add_(0, idx, dZ)
            table.grad = g
            if self.task == "FIT_LATENT":
                log.update(mse_loss=res.mse_loss, prior_loss=res.prior_loss, cosine_loss=res.cosine_loss)
        if need_dw:
            if exchange:
                self.flat.all_reduce_mean(self.group)  # the ONE exchange step of data-parallel training
            self.flat.attach()
        if exchange:
            self._sync_latent_grads()
        self.last_output = res.out
        return log

    def _vad_call(self, name: str, mu, log_var, idx, eps, *rest) -> None:
        import ctypes as C

        from . import _lib

        def arg(a):
            if isinstance(a, torch.Tensor):
                return C.c_void_p(a.data_ptr())
            return a

        with torch.cuda.device(self.device):
            rc = getattr(_lib.load(), name)(arg(mu.detach()), arg(log_var.detach()), arg(idx), arg(eps),
                                            *[arg(a) for a in rest],
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc, name)

    # -- latent tables under data parallelism -------------------------------------------------------------------
    def _latent_params(self):
        return [p for p in (getattr(self.model, n, None) for n in ("Z", "mu", "log_var")) if p is not None]

    def _sync_latent_grads(self) -> None:
        """latent_sync="replicated": sum the dense table gradients over the ranks (each rank's is already divided by the
        world size), as the reference's DDP does for the replicated table."""
        if self.world_size == 1 or self.latent_sync != "replicated":
            return
        for p in self._latent_params():
            if p.grad is not None and p.requires_grad:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
                if not self.ddp_latent_scaling:
                    p.grad.div_(self.world_size)

    def _check_shard(self, idx) -> None:
        """latent_sync="local": a rank may only touch the latents of its own maps.  Checked when the indices are host
        tensors / lists (no device sync); a DistributedSampler that shuffles across ranks trips this."""
        if self.world_size == 1 or self.latent_sync != "local":
            return
        t = torch.as_tensor(idx)
        if t.device.type != "cpu" or t.numel() == 0:
            return
        rank = dist.get_rank(self.group)
        lo, hi = shard_range(self._latent_table().shape[0], rank, self.world_size)
        if int(t.min()) < lo or int(t.max()) >= hi:
            raise ValueError(
                f"rank {rank} owns latent rows [{lo}, {hi}) (shard_range) but the batch indexes "
                f"[{int(t.min())}, {int(t.max())}]: with latent_sync='local' every rank must be fed its own shard of "
                "maps; use latent_sync='replicated' for the reference's replicated-table behaviour")

    @torch.no_grad()
    def gather_latents(self) -> Dict[str, torch.Tensor]:
        """Full latent tables for a checkpoint: with latent_sync="local" every rank holds trained rows only for its own
        shard (shard_range), so the tables are assembled from the owners' rows.  Collective: call on every rank."""
        out = {}
        for name in ("Z", "mu", "log_var"):
            p = getattr(self.model, name, None)
            if p is None:
                continue
            full = p.detach().clone()
            if self.world_size > 1 and self.latent_sync == "local":
                n = full.shape[0]
                rank = dist.get_rank(self.group)
                lo, hi = shard_range(n, rank, self.world_size)
                mine = torch.zeros_like(full)
                mine[lo:hi] = full[lo:hi]
                dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=self.group)  # disjoint shards: the sum assembles them
                full = mine
            out[name] = full
        return out

    def step(self, batch) -> Dict[str, torch.Tensor]:
        """training_step + optimiser step (what trainer.fit does per iteration)."""
        log = self.training_step(batch)
        self.optimizer.step()
        return log
