"""gbrl_b200 -- B200-native (sm_100a) fit/predict engine behind the gbrl `GBRL` surface.

`GBRL` mirrors the reference's pybind11 class gbrl_cpp.GBRL (gbrl/src/cpp/binding.cpp:421-1134) for the
hot path: step / fit / predict with the shared actor-critic tree and the SGD leaf optimizer.
All compute is in libgbrl_b200.so (hand-written CUDA, C-ABI in include/gbrl_b200.h); there is no CPU path.
"""
from .gbrl_cpp import GBRL  # noqa: F401

GBRL_CPP = GBRL   # the alias gbrl/__init__.py exposes for its compiled module

__all__ = ["GBRL", "GBRL_CPP"]
