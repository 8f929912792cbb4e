"""elastic_elgamal_b200 -- B200-native batch engine for elastic-elgamal's verification hot path.

The product is the CUDA library `libeg_b200.so` (C ABI: include/eg_b200.h).  This package is the thin host
mirror used by the tests and the benchmark; it never falls back to a CPU implementation.
"""
from ._ffi import (V_CHALLENGE_MISMATCH, V_CHOICE_RANGE, V_CHOICE_SUM, V_MALFORMED, V_OK, V_QV_CREDIT_EQUIV,
                   V_QV_CREDIT_RANGE, V_QV_VARIANT_BASE)
from .engine import Engine, EngineError

__all__ = ["Engine", "EngineError", "V_OK", "V_MALFORMED", "V_CHALLENGE_MISMATCH", "V_CHOICE_SUM", "V_CHOICE_RANGE",
           "V_QV_CREDIT_RANGE", "V_QV_CREDIT_EQUIV", "V_QV_VARIANT_BASE"]
