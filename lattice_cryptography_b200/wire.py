"""Packed wire format for verification keys and signatures, and a public-seed `key_ch`.

The reference has neither: keys and signatures are live Python objects whose `str()` is a memory address
(`one_time_keys.py:197-237`) and the public row `key_ch` is drawn from `secrets`
(`one_time_keys.py:284-290`), so it has to be shipped as 6.6 / 11.8 KB of coefficients.  Everything in this
module is therefore OPT-IN and does not change what the drop-in API computes:

* `pack_signatures` / `unpack_signatures`: centred coefficients in `sig_bits(pp)` bits (11 at secpar 128,
  13 at 256) instead of 16 - 4,576 / 9,568 bytes per signature instead of 6,656 / 11,776;
* `pack_keys` / `unpack_keys`: NTT-form verification keys in `key_bits(pp)` bits (14 / 16);
* `verify_batch_packed`: `lm_one_time_sigs.verify_batch` straight from the packed records;
* `setup_parameters_from_seed`: `make_setup_parameters` with `key_ch = H2PV('KEY_CH_SALT' || seed)` sampled on
  the GPU with the reference's own distribution parameters (bd = q//2, wt = d), so that two parties derive
  the same row from a short public string.

* `one_time_keys.set_content_str(True)` (re-exported here as `set_content_str`): `str()` / `repr()` of
  verification keys and public statements become content digests, so signatures verify against pickled or
  re-created key objects (changes the challenge bytes, hence every signature; off by default).

Bit layout (include/lcb200.h): value i of a polynomial at bit offset i*bits, least significant bit first.
"""
from math import ceil, log2
from typing import Any, Dict

import numpy as np

from .lattice_algebra import PolynomialVector, engine_for
from .one_time_keys import SchemeParameters, set_content_str  # noqa: F401

KEY_CH_SALT = 'KEY_CH_SALT'


def sig_bits(pp: Dict[str, Any], bound_key: str = 'vf_bd') -> int:
    """Bits per signature coefficient: the range [-bd, bd] has 2*bd + 1 values."""
    return max(1, ceil(log2(2 * pp[bound_key] + 1)))


def key_bits(pp: Dict[str, Any]) -> int:
    return ceil(log2(pp['scheme_parameters'].lp.modulus))


def _engine(pp):
    sp = pp['scheme_parameters']
    return engine_for(sp.lp, sp.secpar)


def pack_signatures(pp, sig, device: bool = False, bound_key: str = 'vf_bd'):
    """sig int16[..., l, 256] -> (uint8[..., l, 32*bits], in_range uint8[..., l]).  A polynomial with a
    coefficient outside [-bd, 2^bits - 1 - bd] cannot be represented; its in_range flag is 0."""
    return _engine(pp).pack(sig, sig_bits(pp, bound_key), pp[bound_key], device=device, want_range=True)


def unpack_signatures(pp, packed, device: bool = False, bound_key: str = 'vf_bd'):
    return _engine(pp).unpack(packed, sig_bits(pp, bound_key), pp[bound_key], dtype=np.int16, device=device)


def pack_keys(pp, vk_ntt, device: bool = False):
    """vk_ntt uint16[..., 256] (engine NTT form, values < q) -> uint8[..., 32*key_bits]."""
    return _engine(pp).pack(vk_ntt, key_bits(pp), 0, device=device)


def unpack_keys(pp, packed, device: bool = False):
    return _engine(pp).unpack(packed, key_bits(pp), 0, dtype=np.uint16, device=device)


def verify_batch_packed(pp, vk_packed, chmsgs, sig_packed, device: bool = False):
    """N verdicts (uint8) from packed keys uint8[N, 2, 32*key_bits] and signatures uint8[N, l, 32*sig_bits]."""
    from .lm_one_time_sigs import _ctx
    eng, sch = _ctx(pp)
    return eng.lm_verify_packed(sch, vk_packed, key_bits(pp), chmsgs, sig_packed, sig_bits(pp), pp['vf_bd'],
                                pp['vf_bd'], pp['vf_wt'], device=device)


def key_ch_from_seed(lp, secpar: int, seed: str) -> PolynomialVector:
    """Public row from a public string: H2PV(KEY_CH_SALT || seed) with bd = q//2, wt = d (the parameters of
    `SchemeParameters.__init__`, one_time_keys.py:284-290, with the hash in place of `secrets`)."""
    coef, _ = engine_for(lp, secpar).hash2polyvec(KEY_CH_SALT, [seed], lp.modulus // 2, lp.degree, lp.length)
    return PolynomialVector(lp, const_time_flag=False, _coef=coef[0])


def setup_parameters_from_seed(make_setup_parameters, secpar: int, seed: str):
    """`make_setup_parameters(secpar)` of any of the three scheme modules, with key_ch derived from `seed`."""
    pp = make_setup_parameters(secpar)
    sp = pp['scheme_parameters']
    pp['scheme_parameters'] = SchemeParameters(secpar=sp.secpar, lp=sp.lp, distribution=sp.distribution,
                                               key_ch=key_ch_from_seed(sp.lp, sp.secpar, seed))
    return pp
