"""CPU tests of the boundary: liblcb200.so loads, exports every symbol include/lcb200.h declares, the
ctypes prototypes cover exactly those symbols, and - with no GPU here - context creation fails
loudly instead of falling back to a CPU path.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'lcb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lcb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from lattice_cryptography_b200 import _ffi
    assert os.path.exists(_ffi.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 30
    for name in syms:
        assert hasattr(lib, name), f'{name} declared in include/lcb200.h but not exported'
    assert sorted(_ffi.PROTOTYPES) == syms, 'ctypes prototypes and the header disagree'


def test_scheme_struct_layout_matches_header():
    from lattice_cryptography_b200 import _ffi, make_scheme
    assert ctypes.sizeof(_ffi.LcbScheme) == 8 * 4 + 4 * _ffi.LCB_SALT_MAX
    s = make_scheme(sk_bd=45, sk_wt=256, ch_bd=1, ch_wt=20)
    assert (s.sk_bd, s.sk_wt, s.ch_wt, s.sk_salt, s.ag_salt) == (45, 256, 20, b'SK_SALT', b'AG_SALT')
    with pytest.raises(ValueError):
        make_scheme(sk_salt='x' * 40)


def test_no_cpu_fallback():
    import torch
    from lattice_cryptography_b200 import Engine, LcbError, _ffi
    lib = _ffi.load()
    assert lib.lcb_version() >= 100
    assert b'no CPU fallback' in lib.lcb_strerror(_ffi.LCB_ERR_NO_DEVICE)
    if torch.cuda.is_available():
        pytest.skip('a GPU is present; the fallback check is for CPU-only boxes')
    with pytest.raises(LcbError) as e:
        Engine(128, 11777, 256, 13)
    assert e.value.status == _ffi.LCB_ERR_NO_DEVICE
    # unsupported parameter sets are rejected before any device work
    for bad in ((128, 11777, 16, 13), (128, 11777, 2048, 13), (128, 11777, 96, 13), (128, 11779, 256, 13),
                (128, 193, 256, 1), (128, 11777, 256, 0), (128, 11777, 256, 65), (0, 11777, 256, 13), (513, 11777, 256, 13)):
        with pytest.raises(LcbError) as e:
            Engine(*bad)
        assert e.value.status == _ffi.LCB_ERR_INVALID


def test_ragged_packing():
    from lattice_cryptography_b200 import ragged
    blob, off = ragged(['ab', b'', 'cé'])
    assert off.tolist() == [0, 2, 2, 5] and bytes(blob) == b'abc\xc3\xa9'
    blob, off = ragged([])
    assert off.tolist() == [0] and blob.size == 0


def test_dropin_modules_import_without_gpu():
    """Importing the drop-in modules must not touch the GPU (parameter tables are host logic)."""
    from lattice_cryptography_b200 import adaptor_sigs, bklm_one_time_agg_sigs, lm_one_time_sigs, one_time_keys
    assert lm_one_time_sigs.LPs[128].rou == 24 and lm_one_time_sigs.LPs[256].rou == 55
    assert bklm_one_time_agg_sigs.CAPs == {128: 2, 256: 2}
    assert adaptor_sigs.WTs[128]['wit_wt'] == 20
    assert lm_one_time_sigs.distribute_tasks(list(range(10)), 4) == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert one_time_keys.bits_to_indices is one_time_keys.bits_per_index_set
