"""B200-native (sm_100a) fused window-attention super-resolution: drop-in modules for the reference's
SwinIR blocks (ViacheslavTimofeev/tpu_superresolution, modules/network_swinir.py) backed by libsrk.so."""
from .swinir import (Mlp, WindowAttention, SwinTransformerBlock, BasicLayer, RSTB, PatchEmbed, PatchUnEmbed,
                     PixelShuffle, Upsample, UpsampleOneStep, SwinIR, calculate_mask)

__all__ = ["Mlp", "WindowAttention", "SwinTransformerBlock", "BasicLayer", "RSTB", "PatchEmbed", "PatchUnEmbed",
           "PixelShuffle", "Upsample", "UpsampleOneStep", "SwinIR", "calculate_mask"]
