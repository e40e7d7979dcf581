"""CPU ORACLE (test infrastructure): loads the reference's OWN scheme modules, unmodified,
from /root/reference on top of the restated ``lattice_algebra`` in this directory.

This only works in the build container (the GPU box has no /root/reference), so it is used
by oracle/gen_golden.py to produce the committed fixtures under tests/golden/ and by
`-m "not gpu"` tests that cross-check oracle/schemes.py against the reference when the
reference is present.  Nothing on the product path imports this.

Two in-process fixes are applied, neither touching /root/reference:
  * `lattice_algebra` resolves to oracle/lattice_algebra (the real 0.1.1 package is absent);
  * reference HEAD imports `bits_to_indices`/`bits_to_decode` from one_time_keys
    (lm_one_time_sigs.py:4-5, bklm_one_time_agg_sigs.py:5, adaptor_sigs.py:2-3) but
    one_time_keys.py:243-256 defines them as `bits_per_index_set`/`bits_per_coefficient`;
    the two names are aliased before the dependants are imported.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('LCB_REFERENCE_ROOT', '/root/reference')
_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'lattice_cryptography'))


def load_reference():
    """Returns (one_time_keys, lm_one_time_sigs, bklm_one_time_agg_sigs, adaptor_sigs) of the reference."""
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    if _ORACLE_DIR not in sys.path:
        sys.path.insert(0, _ORACLE_DIR)          # makes `import lattice_algebra` hit the restatement
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    otk = importlib.import_module('lattice_cryptography.one_time_keys')
    if not hasattr(otk, 'bits_to_indices'):
        otk.bits_to_indices = otk.bits_per_index_set
    if not hasattr(otk, 'bits_to_decode'):
        otk.bits_to_decode = otk.bits_per_coefficient
    lm = importlib.import_module('lattice_cryptography.lm_one_time_sigs')
    bklm = importlib.import_module('lattice_cryptography.bklm_one_time_agg_sigs')
    ad = importlib.import_module('lattice_cryptography.adaptor_sigs')
    return otk, lm, bklm, ad
