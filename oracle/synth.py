"""Synthetic weights / inputs moved to ``tpu_superresolution_b200/synth.py`` (data generation is not part of the checker);
this module re-exports them so fixtures and tests keep one import path."""
from tpu_superresolution_b200.synth import *  # noqa: F401,F403
from tpu_superresolution_b200.synth import _draw  # noqa: F401
