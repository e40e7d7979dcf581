"""lattice_cryptography_b200 — B200-native batched engine for the hot path of
b-g-goodell/lattice-cryptography (LM one-time signatures, BKLM aggregation, adaptor signatures).

`Engine` is the array-level face of the C ABI (include/lcb200.h); the modules
`lm_one_time_sigs`, `bklm_one_time_agg_sigs`, `adaptor_sigs`, `one_time_keys` and
`lattice_algebra` keep the reference's Python entry points and add `*_batch` variants.
There is no CPU fallback.
"""
__version__ = '0.1.0'

from ._ffi import LcbError, LcbScheme  # noqa: F401
from .engine import Engine, MultiEngine, make_scheme, ragged  # noqa: F401
