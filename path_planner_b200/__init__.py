"""path_planner_b200 -- B200-native batched Dubins edge-evaluation engine.

The hot path of afb2001/path_planner (Edge::computeTrueCost and the Dubins solve in front of it)
as hand-written sm_100a CUDA kernels behind the C ABI of include/ppe.h.  Importing the package
does not load the CUDA library; constructing an `EdgeEngine` does, and fails loudly when the
library or the GPU is missing (there is no CPU fallback).
"""
from . import abi, mapio, synth  # noqa: F401  (sharding imports torch: import it explicitly)
from ._capi import PpeError  # noqa: F401
from .engine import EdgeEngine, load_library, LIB_PATH  # noqa: F401

__all__ = ["abi", "mapio", "synth", "EdgeEngine", "PpeError", "load_library", "LIB_PATH"]
