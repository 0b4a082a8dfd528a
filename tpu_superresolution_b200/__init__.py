"""B200-native (sm_100a) fused window-attention super-resolution: drop-in modules for the reference's
SwinIR / HAT / DAT blocks (ViacheslavTimofeev/tpu_superresolution, modules/network_swinir.py, hat_arch.py, dat_arch.py) backed by libsrk.so."""
from .swinir import (Mlp, WindowAttention, SwinTransformerBlock, BasicLayer, RSTB, PatchEmbed, PatchUnEmbed,
                     PixelShuffle, Upsample, UpsampleOneStep, SwinIR, calculate_mask, set_precision)
from . import hat
from .hat import HAT, HAB, OCAB, RHAG, CAB, ChannelAttention, AttenBlocks
from . import dat
from .dat import DAT, DATB, ResidualGroup, Adaptive_Spatial_Attention, Adaptive_Channel_Attention, SGFN, SpatialGate, DynamicPosBias
from .graphs import GraphedModel, PipelinedRunner
from .convs import invalidate_all as invalidate_packed      # for patched reference containers (models have .invalidate_packed())

__all__ = ["Mlp", "WindowAttention", "SwinTransformerBlock", "BasicLayer", "RSTB", "PatchEmbed", "PatchUnEmbed",
           "PixelShuffle", "Upsample", "UpsampleOneStep", "SwinIR", "calculate_mask",
           "hat", "HAT", "HAB", "OCAB", "RHAG", "CAB", "ChannelAttention", "AttenBlocks",
           "dat", "DAT", "DATB", "ResidualGroup", "Adaptive_Spatial_Attention", "Adaptive_Channel_Attention", "SGFN", "SpatialGate",
           "DynamicPosBias", "GraphedModel", "PipelinedRunner", "set_precision", "invalidate_packed"]
