"""Differentiable torch forms of the reference's losses, under the reference's names.

Reference: src/utils/loss_functions.py:6-71.  They exist for callers that evaluate the loss on the
decoder output themselves -- ``loss.backward()`` then reaches the fused kernels through
``reni_b200.functional.decode``.  The fully fused step (forward + loss + backward inside the CUDA
library, no (B, P, 3) autograd graph) is ``reni_b200.training.training_step``.

Semantics worth spelling out because the fused kernels reproduce them exactly:
  * WeightedMSE is a MEAN over pixels x channels but a SUM over the maps of the batch;
  * WeightedCosineSimilarity takes the cosine over the PIXEL axis (dim=1), one value per
    (map, channel), multiplies it by the sine weight of pixel 0 only, averages over channels and
    sums ``1 - .`` over the batch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F


def _per_map(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(x.shape[0], -1)


def WeightedMSE(model_output, ground_truth, sineweight):
    """loss_functions.py:6-13"""
    err = (model_output - ground_truth).square() * sineweight
    return _per_map(err).mean(dim=1).sum()


def KLD(mu, log_var, Z_dims=1):
    """loss_functions.py:16-22"""
    per_map = _per_map(1 + log_var - mu.square() - log_var.exp()).sum(dim=1)
    return (-0.5 * per_map / Z_dims).sum()


def WeightedCosineSimilarity(model_output, ground_truth, sineweight):
    """loss_functions.py:25-32"""
    cs = F.cosine_similarity(model_output, ground_truth, dim=1, eps=1e-20)  # (B, 3): over pixels
    return (1 - (cs * sineweight[:, 0]).mean(dim=1)).sum()


def CosineSimilarity(model_output, ground_truth):
    """loss_functions.py:35-36"""
    return 1 - F.cosine_similarity(model_output, ground_truth, dim=1, eps=1e-20).mean()


class RENITrainLoss:
    """FIT_DECODER / AutoDecoder criterion (loss_functions.py:39-45)."""

    def __call__(self, inputs, targets, sineweight):
        return WeightedMSE(inputs, targets, sineweight)


@dataclass
class RENIVADTrainLoss:
    """FIT_DECODER / VariationalAutoDecoder criterion (loss_functions.py:47-58) -> (loss, mse, kld)."""

    beta: float = 1
    Z_dims: Optional[int] = None

    def __call__(self, inputs, targets, sineweight, mu, log_var):
        mse = WeightedMSE(inputs, targets, sineweight)
        kl = self.beta * KLD(mu, log_var, self.Z_dims)
        return mse + kl, mse, kl


@dataclass
class RENITestLoss:
    """FIT_LATENT criterion (loss_functions.py:60-71) -> (loss, mse, prior, cosine)."""

    alpha: float = 1
    beta: float = 1

    def __call__(self, inputs, targets, sineweight, Z):
        mse = WeightedMSE(inputs, targets, sineweight)
        prior = self.alpha * Z.square().sum()
        cosine = self.beta * WeightedCosineSimilarity(inputs, targets, sineweight)
        return mse + prior + cosine, mse, prior, cosine
