"""reni_b200 -- B200-native (sm_100a) implementation of the RENI decoder hot path.

Drop-in for the Cond-by-Concat and FiLM decoder families of JADGardner/RENI (src/models/RENI.py) and the
training / latent-optimisation step around it (src/lightning/RENI_module.py:80-146):

    from reni_b200 import RENIAutoDecoder, RENIVariationalAutoDecoder, get_model   # same ctor args
    from reni_b200 import RENITrainLoss, RENITestLoss, get_directions, get_sineweight
    from reni_b200 import RENITrainer                                              # fused step + data parallel

All decoder arithmetic runs in the hand-written CUDA library ``reni_b200/lib/libreni_b200.so``
(C ABI: include/reni_b200.h), built by ``__graft_entry__.build()``.  There is no CPU fallback.
"""
from .geometry import get_directions, get_mask, get_sineweight, pack_mask_bits, rectangle_mask
from .losses import (KLD, CosineSimilarity, RENITestLoss, RENITrainLoss, RENIVADTrainLoss, WeightedCosineSimilarity,
                     WeightedMSE)
from .film import CustomMappingNetwork, FiLMLayer, RENIAutoDecoderFiLM, RENIVariationalAutoDecoderFiLM
from .models import RENIAutoDecoder, RENIVariationalAutoDecoder, SineLayer, get_model
from .optim import FusedAdam
from .render import EnvironmentMap, blinn_phong_shading_env_map
from .serving import GraphedDecoder
from .training import FlatGradBuffer, RENITrainer, shard_range

__all__ = [
    "RENIAutoDecoder", "RENIVariationalAutoDecoder", "SineLayer", "get_model",
    "RENIAutoDecoderFiLM", "RENIVariationalAutoDecoderFiLM", "FiLMLayer", "CustomMappingNetwork",
    "WeightedMSE", "KLD", "WeightedCosineSimilarity", "CosineSimilarity",
    "RENITrainLoss", "RENIVADTrainLoss", "RENITestLoss",
    "get_directions", "get_sineweight", "get_mask", "rectangle_mask", "pack_mask_bits",
    "RENITrainer", "FlatGradBuffer", "shard_range", "FusedAdam", "GraphedDecoder",
    "EnvironmentMap", "blinn_phong_shading_env_map",
]
