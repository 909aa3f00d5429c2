"""rscm_b200 — B200-native ensemble engine for the RSCM hot path.

Submodules mirror the reference's pyo3 layout ``rscm._lib.{core,components,two_layer,magicc,calibrate}``
(crates/rscm/src/python/mod.rs:47-66) for the subset that runs on the GPU.
Importing the package requires the built CUDA library (no CPU fallback).
"""

from . import _ffi  # noqa: F401  (fails loudly when librscm_b200.so is missing)
from . import calibrate, components, core, magicc, two_layer  # noqa: F401

__version__ = "0.1.0"
