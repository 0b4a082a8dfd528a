"""B200-native (sm_100a) fused window-attention super-resolution: drop-in modules for the reference's
SwinIR / HAT blocks (ViacheslavTimofeev/tpu_superresolution, modules/network_swinir.py, modules/hat_arch.py) backed by libsrk.so."""
from .swinir import (Mlp, WindowAttention, SwinTransformerBlock, BasicLayer, RSTB, PatchEmbed, PatchUnEmbed,
                     PixelShuffle, Upsample, UpsampleOneStep, SwinIR, calculate_mask)
from . import hat
from .hat import HAT, HAB, OCAB, RHAG, CAB, ChannelAttention, AttenBlocks

__all__ = ["Mlp", "WindowAttention", "SwinTransformerBlock", "BasicLayer", "RSTB", "PatchEmbed", "PatchUnEmbed",
           "PixelShuffle", "Upsample", "UpsampleOneStep", "SwinIR", "calculate_mask",
           "hat", "HAT", "HAB", "OCAB", "RHAG", "CAB", "ChannelAttention", "AttenBlocks"]
