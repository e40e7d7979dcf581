"""CPU ORACLE (test infrastructure): ctypes face of oracle/lcb_oracle.c (schoolbook / bit-by-bit C
restatement).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this."""
import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_int, c_int64, c_long, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'liblcb_oracle.so')


class OrcParams(Structure):
    _fields_ = [('secpar', c_int), ('q', c_int), ('d', c_int), ('l', c_int), ('sk_bd', c_int), ('sk_wt', c_int),
                ('ch_bd', c_int), ('ch_wt', c_int), ('sk_salt', c_char_p), ('ch_salt', c_char_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(['make', '-C', _HERE], check=True, capture_output=True)
        _lib = ctypes.CDLL(_PATH)
        _lib.orc_omp_threads.restype = c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p) if a is not None else None


def params(secpar, q, l, sk_bd, ch_wt, d=256, sk_wt=256, ch_bd=1, sk_salt=b'SK_SALT', ch_salt=b'CH_SALT'):
    return OrcParams(secpar, q, d, l, sk_bd, sk_wt, ch_bd, ch_wt, sk_salt, ch_salt)


def shake256(data: bytes, n: int) -> bytes:
    out = np.empty(n, dtype=np.uint8)
    lib().orc_shake256(data, c_size_t(len(data)), _p(out), c_size_t(n))
    return bytes(out)


def hash2polyvec(secpar, d, salt: str, msg: bytes, bd, wt, vec_len):
    dense = np.empty((vec_len, d), dtype=np.int16)
    pairs = np.empty((vec_len, wt, 2), dtype=np.int16)
    rc = lib().orc_hash2polyvec(secpar, d, salt.encode(), msg, c_size_t(len(msg)), bd, wt, vec_len, _p(dense), _p(pairs))
    assert rc == 0
    return dense, pairs


def agg_coef(secpar, d, ag_salt: str, index: int, agmsg: bytes):
    out = np.empty(2, dtype=np.int16)
    assert lib().orc_agg_coef(secpar, d, ag_salt.encode(), c_long(index), agmsg, c_size_t(len(agmsg)), _p(out)) == 0
    return int(out[0]), int(out[1])


def poly_mul(q, a, b):
    out = np.empty_like(a)
    lib().orc_poly_mul(q, a.shape[-1], _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out


def dot(q, row, vec):
    """sum_i row[i] * vec[i] (PolynomialVector.__mul__), int16[l][d] x int16[l][d] -> int16[d]"""
    out = np.empty(row.shape[-1], dtype=np.int16)
    lib().orc_dot(q, row.shape[-1], row.shape[0], _p(np.ascontiguousarray(row)), _p(np.ascontiguousarray(vec)), _p(out))
    return out


def lm_keygen(p: OrcParams, key_ch, seed: bytes):
    skl = np.empty((p.l, p.d), dtype=np.int16)
    skr = np.empty((p.l, p.d), dtype=np.int16)
    vkl = np.empty(p.d, dtype=np.int16)
    vkr = np.empty(p.d, dtype=np.int16)
    rc = lib().orc_lm_keygen(ctypes.byref(p), _p(key_ch), seed, c_size_t(len(seed)), _p(skl), _p(skr), _p(vkl), _p(vkr))
    assert rc == 0
    return skl, skr, vkl, vkr


def lm_sign(p: OrcParams, skl, skr, chmsg: bytes):
    sig = np.empty((p.l, p.d), dtype=np.int16)
    assert lib().orc_lm_sign(ctypes.byref(p), _p(skl), _p(skr), chmsg, c_size_t(len(chmsg)), _p(sig)) == 0
    return sig


def lm_verify(p: OrcParams, key_ch, vkl, vkr, chmsg: bytes, sig, bd, wt, st=None) -> bool:
    r = lib().orc_lm_verify(ctypes.byref(p), _p(key_ch), _p(np.ascontiguousarray(vkl)), _p(np.ascontiguousarray(vkr)),
                            chmsg, c_size_t(len(chmsg)), _p(np.ascontiguousarray(sig)),
                            _p(np.ascontiguousarray(st)) if st is not None else None, bd, wt)
    assert r >= 0
    return bool(r)


def lm_verify_batch(p: OrcParams, key_ch, vk_coef, blob, off, sig, bd, wt):
    n = int(off.shape[0]) - 1
    verdict = np.empty(n, dtype=np.uint8)
    rc = lib().orc_lm_verify_batch(ctypes.byref(p), _p(key_ch), _p(np.ascontiguousarray(vk_coef)),
                                   _p(np.ascontiguousarray(blob)), _p(np.ascontiguousarray(off.astype(np.int64))),
                                   _p(np.ascontiguousarray(sig)), c_long(n), bd, wt, _p(verdict))
    assert rc == 0
    return verdict


def omp_threads() -> int:
    return lib().orc_omp_threads()
