"""CPU ORACLE (test infrastructure): loads the reference's OWN scheme modules, unmodified,
from /root/reference on top of the restated ``lattice_algebra`` in this directory.

This only works in the build container (the GPU box has no /root/reference), so it is used
by oracle/gen_golden.py to produce the committed fixtures under tests/golden/ and by
`-m "not gpu"` tests that cross-check oracle/schemes.py against the reference when the
reference is present.  Nothing on the product path imports this.

Two in-process fixes are applied, neither touching /root/reference:
  * `lattice_algebra` resolves to the REAL package when one is reachable (oracle/l1.py: $LCB_LATTICE_ALGEBRA,
    baseline/_ref, site-packages) and to the restatement oracle/lattice_algebra otherwise (the real 0.1.1 package
    is absent from this image); `L1_LABEL` says which ("restated" | "lattice_algebra==<version>");
  * reference HEAD imports `bits_to_indices`/`bits_to_decode` from one_time_keys
    (lm_one_time_sigs.py:4-5, bklm_one_time_agg_sigs.py:5, adaptor_sigs.py:2-3) but
    one_time_keys.py:243-256 defines them as `bits_per_index_set`/`bits_per_coefficient`;
    the two names are aliased before the dependants are imported.
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))) if os.path.dirname(os.path.abspath(__file__)) not in sys.path else None

REFERENCE_ROOT = os.environ.get('LCB_REFERENCE_ROOT', '/root/reference')
_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
L1_LABEL = None        # set by load_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'lattice_cryptography'))


def load_reference():
    """Returns (one_time_keys, lm_one_time_sigs, bklm_one_time_agg_sigs, adaptor_sigs) of the reference."""
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    global L1_LABEL
    import l1 as l1_select
    real = l1_select.find_real()
    L1_LABEL = l1_select.label_of(real)
    if 'lattice_algebra' in sys.modules:
        loaded = os.path.dirname(os.path.dirname(os.path.abspath(sys.modules['lattice_algebra'].__file__)))
        if real is not None and loaded != real:
            raise RuntimeError('a real lattice_algebra is reachable but the restatement was imported first; '
                               'import ref_loader.load_reference() before anything imports lattice_algebra')
    if _ORACLE_DIR not in sys.path:
        sys.path.insert(0, _ORACLE_DIR)          # makes `import lattice_algebra` hit the restatement ...
    if real is not None:
        sys.path.insert(0, real)                 # ... unless the real package exists: it goes first
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
