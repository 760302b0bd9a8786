"""hystrath_b200 -- B200-native engine for dsmcFoam+'s per-timestep particle loop
(dsmcCloud::evolve): move -> sort -> NTC select -> VHS/LB collide -> sample.

The product is libdsmcb200.so (hand-written sm_100a CUDA behind the C ABI of
include/dsmcb200.h); this package is its host-side mirror of the reference interface.
"""
from . import capi, meshgen  # noqa: F401

__all__ = ["capi", "meshgen"]
