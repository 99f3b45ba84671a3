"""Fixed-shape, no-grad decoding through one CUDA-graph replay.

Decoding a handful of latents (rendering one environment map from its latent code, the notebook's and the callbacks'
use of the model: examples.ipynb, src/lightning/callbacks.py:40-43,93) is host-bound when issued op by op -- about
0.47 ms for a single 64x128 FiLM decode on a B200 whose kernels need a small fraction of that.  ``GraphedDecoder``
captures ``model(Z, directions)`` once for a fixed batch and direction grid (per-map stage, fp32 -> fp16 weight images
and the fused decoder kernel) and replays it; the decoder weights are read at replay time, so an optimiser step or a
``load_state_dict`` between calls is picked up.
"""
from __future__ import annotations

import torch

from .film import _FilmDecoderBase
from .functional import Workspace


class GraphedDecoder:
    def __init__(self, model: torch.nn.Module, batch: int, directions: torch.Tensor):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("reni_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
        self.model = model
        self.Z = torch.zeros(batch, model.ndims, 3, device=dev)
        self.directions = directions.to(dev)
        if self.directions.shape[0] not in (1, batch):
            raise ValueError(f"directions must have batch 1 or {batch}, got {self.directions.shape[0]}")
        if self.directions.shape[0] == 1 and batch > 1:
            self.directions = self.directions.expand(batch, -1, -1)  # one grid shared by all maps (stride 0)
        self._film = isinstance(model, _FilmDecoderBase)
        # the captured graph holds raw pointers into its workspace, so it owns one: the model's own inference workspace
        # may be replaced (grown) by a later eager call with a larger batch
        self._ws = Workspace()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():  # warm-up outside capture: workspace, lazy initialisation
            for _ in range(2):
                self._decode()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._ws.prepared_key = None  # rebuild the weight images INSIDE the graph: replays follow weight updates
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self._decode()

    def _decode(self) -> torch.Tensor:
        saved_ws, self.model._ws = self.model._ws, self._ws
        try:
            return self._decode_inner()
        finally:
            self.model._ws = saved_ws

    def _decode_inner(self) -> torch.Tensor:
        if self._film:  # the one-launch native per-map stage (reni_film_map_forward) instead of ~25 torch launches
            saved = self.model.NATIVE_MAP_LEVEL_MAX_BATCH
            self.model.NATIVE_MAP_LEVEL_MAX_BATCH = max(saved, self.Z.shape[0])
            try:
                return self.model(self.Z, self.directions)
            finally:
                self.model.NATIVE_MAP_LEVEL_MAX_BATCH = saved
        return self.model(self.Z, self.directions)

    @torch.no_grad()
    def __call__(self, Z: torch.Tensor) -> torch.Tensor:
        """Z (batch, N, 3) latent codes -> (batch, P, out_features); the returned tensor is overwritten by the next call."""
        self.Z.copy_(Z, non_blocking=True)
        self.graph.replay()
        return self.out
