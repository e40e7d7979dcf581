"""CPU ORACLE (test infrastructure, never the product path): restatement of the reference's
scheme layer with EXPLICIT hash-input strings, on top of oracle/lattice_algebra.

PARITY UNPINNED at the byte level (see oracle/lattice_algebra/__init__.py); this file's
scheme logic is cross-checked against the reference's own modules, imported unmodified by
oracle/ref_loader.py, in tests/test_oracle.py (runs wherever /root/reference exists) and
through the fixtures in tests/golden/ (made by oracle/gen_golden.py from the reference's
modules).

Why explicit strings: the reference hashes `str(otvk) + ', ' + msg`
(lm_one_time_sigs.py:148) and `str(st) + ', ' + str(otvk) + ', ' + msg`
(adaptor_sigs.py:176) where str() of those key objects is the CPython default
`<... object at 0x...>`; the aggregation message is `str(list(zip(keys, msgs)))`
(bklm_one_time_agg_sigs.py:65).  The engine's C ABI therefore takes the full hash input as
bytes, and so does this oracle.

Function bodies follow, line for line in meaning:
  make_lm_parameters / make_bklm_parameters / make_adaptor_parameters
        lm_one_time_sigs.py:36-55, bklm_one_time_agg_sigs.py:27-44, adaptor_sigs.py:37-71
  sample_vector                 lm_one_time_sigs.py:70-91, adaptor_sigs.py:86-96
  lm_keygen_one                 lm_one_time_sigs.py:64-97
  challenge                     lm_one_time_sigs.py:141-160
  lm_sign / lm_verify           lm_one_time_sigs.py:163-191 (adaptor variants adaptor_sigs.py:191-217,247-266)
  agg_coefs / aggregate / aggregate_verify    bklm_one_time_agg_sigs.py:78-81,92-116
  witgen_one / adapt / extract / witness_verify   adaptor_sigs.py:80-101,220-237
"""
import os
import sys
from typing import Dict, List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from lattice_algebra import (LatticeParameters, Polynomial, PolynomialVector, hash2polynomial,  # noqa: E402
                             hash2polynomialvector, bits_to_indices, bits_to_decode,
                             UNIFORM_INFINITY_WEIGHT)

DIST = UNIFORM_INFINITY_WEIGHT

# shipped parameter sets: lm_one_time_sigs.py:19-33, bklm_one_time_agg_sigs.py:15-24, adaptor_sigs.py:19-34
SHIPPED = {
    128: dict(modulus=11777, degree=256, length=13, sk_bd=45, sk_wt=256, ch_bd=1, ch_wt=20,
              ag_bd=1, ag_wt=1, ag_cap=2, wit_bd=1, wit_wt=20),
    256: dict(modulus=39937, degree=256, length=23, sk_bd=65, sk_wt=256, ch_bd=1, ch_wt=50,
              ag_bd=1, ag_wt=1, ag_cap=2, wit_bd=1, wit_wt=20),
}
_LP_CACHE: Dict[Tuple[int, int, int], LatticeParameters] = {}


def lattice_parameters(modulus: int, degree: int, length: int) -> LatticeParameters:
    key = (modulus, degree, length)
    if key not in _LP_CACHE:
        _LP_CACHE[key] = LatticeParameters(modulus=modulus, degree=degree, length=length)
    return _LP_CACHE[key]


# ------------------------------------------------------------------ dense <-> object helpers
def poly_from_dense(lp: LatticeParameters, dense: Sequence[int]) -> Polynomial:
    return Polynomial(lp=lp, coefs={i: int(v) for i, v in enumerate(dense) if int(v) != 0})


def vec_from_dense(lp: LatticeParameters, rows: Sequence[Sequence[int]]) -> PolynomialVector:
    return PolynomialVector(lp=lp, entries=[poly_from_dense(lp, r) for r in rows])


def dense_of_poly(p: Polynomial) -> List[int]:
    out = [0] * p.lp.degree
    for i, v in p.get_coef_rep()[0].items():
        out[i] = v
    return out


def dense_of_vec(v: PolynomialVector) -> List[List[int]]:
    return [dense_of_poly(p) for p in v.entries]


# ------------------------------------------------------------------ parameters
def make_lm_parameters(secpar: int, key_ch: PolynomialVector) -> dict:
    s = SHIPPED[secpar]
    lp = lattice_parameters(s['modulus'], s['degree'], s['length'])
    pp = dict(secpar=secpar, lp=lp, key_ch=key_ch, sk_salt='SK_SALT', ch_salt='CH_SALT',
              sk_bd=s['sk_bd'], sk_wt=s['sk_wt'], ch_bd=s['ch_bd'], ch_wt=s['ch_wt'])
    pp['vf_wt'] = max(1, min(lp.degree, pp['sk_wt'] * (1 + pp['ch_wt'])))
    pp['vf_bd'] = max(1, min(lp.modulus // 2, pp['sk_bd'] * (1 + min(pp['sk_wt'], pp['ch_wt']) * pp['ch_bd'])))
    return pp


def make_bklm_parameters(secpar: int, key_ch: PolynomialVector, ag_cap: Optional[int] = None) -> dict:
    pp = make_lm_parameters(secpar, key_ch)
    s = SHIPPED[secpar]
    lp = pp['lp']
    pp['ag_cap'] = s['ag_cap'] if ag_cap is None else ag_cap
    pp['ag_salt'] = 'AG_SALT'
    pp['ag_bd'] = s['ag_bd']
    pp['ag_wt'] = s['ag_wt']
    pp['avf_wt'] = max(1, min(lp.degree, pp['ag_cap'] * pp['ag_wt'] * pp['vf_wt']))
    pp['avf_bd'] = max(1, min(lp.modulus // 2,
                              pp['ag_cap'] * min(pp['ag_wt'], pp['vf_wt']) * pp['ag_bd'] * pp['vf_bd']))
    return pp


def make_adaptor_parameters(secpar: int, key_ch: PolynomialVector) -> dict:
    s = SHIPPED[secpar]
    lp = lattice_parameters(s['modulus'], s['degree'], s['length'])
    d, half = lp.degree, (lp.modulus - 1) // 2
    pp = dict(secpar=secpar, lp=lp, key_ch=key_ch, sk_salt='SK_SALT', ch_salt='CH_SALT', wit_salt='WIT_SALT',
              sk_bd=s['sk_bd'], sk_wt=min(d, s['sk_wt']), ch_bd=s['ch_bd'], ch_wt=min(d, s['ch_wt']),
              wit_bd=s['wit_bd'], wit_wt=min(d, s['wit_wt']))
    base_bd = pp['sk_bd'] * (1 + min(d, pp['sk_wt'], pp['ch_wt']) * pp['ch_bd'])
    pp['pvf_wt'] = max(1, min(d, pp['sk_wt'] * (1 + pp['ch_wt'])))
    pp['pvf_bd'] = max(1, min(half, base_bd))
    pp['vf_wt'] = max(1, min(d, pp['sk_wt'] * (1 + pp['ch_wt']) + pp['wit_wt']))
    pp['vf_bd'] = max(1, min(half, base_bd + pp['wit_bd']))
    pp['ext_wit_wt'] = max(1, min(d, pp['vf_wt'] + pp['pvf_wt']))
    pp['ext_wit_bd'] = max(1, min(half, pp['vf_bd'] + pp['pvf_bd']))
    return pp


def key_ch_from_seed(secpar: int, seed: str) -> PolynomialVector:
    """Deterministic public row for tests and benches: the same decoder the reference feeds
    with `secrets` (one_time_keys.py:284-290: bd = q//2, wt = d), fed here by SHAKE256 over a
    public seed string.  NOT a reference behaviour (the reference's key_ch is not
    reproducible); it only fixes key_ch so that fixtures are deterministic."""
    s = SHIPPED[secpar]
    lp = lattice_parameters(s['modulus'], s['degree'], s['length'])
    return sample_vector(secpar, lp, 'KEY_CH_SEED', seed, lp.modulus // 2, lp.degree)


# ------------------------------------------------------------------ samplers
def sample_vector(secpar: int, lp: LatticeParameters, salt: str, msg: str, bd: int, wt: int) -> PolynomialVector:
    return hash2polynomialvector(
        secpar=secpar, lp=lp, distribution=DIST, dist_pars={'bd': bd, 'wt': wt}, num_coefs=wt,
        bti=bits_to_indices(secpar=secpar, degree=lp.degree, wt=wt), btd=bits_to_decode(secpar=secpar, bd=bd),
        salt=salt, msg=msg, const_time_flag=True)


def sample_poly(secpar: int, lp: LatticeParameters, salt: str, msg: str, bd: int, wt: int) -> Polynomial:
    return hash2polynomial(
        secpar=secpar, lp=lp, distribution=DIST, dist_pars={'bd': bd, 'wt': wt}, salt=salt, msg=msg, num_coefs=wt,
        bti=bits_to_indices(secpar=secpar, degree=lp.degree, wt=wt), btd=bits_to_decode(secpar=secpar, bd=bd),
        const_time_flag=True)


# ------------------------------------------------------------------ LM one-time signatures
def lm_keygen_one(pp: dict, seed: str):
    """-> (sk_left, sk_right, vk_left, vk_right)"""
    lp, secpar = pp['lp'], pp['secpar']
    left = sample_vector(secpar, lp, pp['sk_salt'] + 'LEFT', seed, pp['sk_bd'], pp['sk_wt'])
    right = sample_vector(secpar, lp, pp['sk_salt'] + 'RIGHT', seed, pp['sk_bd'], pp['sk_wt'])
    return left, right, pp['key_ch'] * left, pp['key_ch'] * right


def challenge(pp: dict, chmsg: str) -> Polynomial:
    return sample_poly(pp['secpar'], pp['lp'], pp['ch_salt'], chmsg, pp['ch_bd'], pp['ch_wt'])


def lm_sign(pp: dict, sk_left: PolynomialVector, sk_right: PolynomialVector, chmsg: str) -> PolynomialVector:
    return sk_left ** challenge(pp, chmsg) + sk_right


def lm_verify(pp: dict, vk_left: Polynomial, vk_right: Polynomial, chmsg: str, sig: PolynomialVector,
              st: Optional[Polynomial] = None, bd_name: str = 'vf_bd', wt_name: str = 'vf_wt') -> bool:
    """LM verify (st None) and the adaptor preverify / verify variants (bd/wt names, + st)."""
    cnws = sig.get_coef_rep()
    n, w = max(i[1] for i in cnws), max(i[2] for i in cnws)
    if n > pp[bd_name] or w > pp[wt_name]:
        return False
    c = challenge(pp, chmsg)
    lhs = pp['key_ch'] * sig
    rhs = vk_left * c + vk_right
    if st is not None:
        rhs = rhs + st
    return lhs == rhs


# ------------------------------------------------------------------ BKLM aggregation
def agg_coefs(pp: dict, agmsg: str, n: int) -> List[Polynomial]:
    return [sample_poly(pp['secpar'], pp['lp'], pp['ag_salt'] + str(i), agmsg, pp['ag_bd'], pp['ag_wt'])
            for i in range(n)]


def aggregate(pp: dict, sorted_sigs: List[PolynomialVector], agmsg: str) -> PolynomialVector:
    coefs = agg_coefs(pp, agmsg, len(sorted_sigs))
    return sum([sig ** a for sig, a in zip(sorted_sigs, coefs)])


def aggregate_verify(pp: dict, sorted_vks: List[Tuple[Polynomial, Polynomial]], sorted_chmsgs: List[str],
                     agmsg: str, ag_sig: PolynomialVector) -> bool:
    cnw = ag_sig.get_coef_rep()
    n, w = max(i[1] for i in cnw), max(i[2] for i in cnw)
    count = len(sorted_vks)
    if n < 1 or n > pp['avf_bd'] or w < 1 or w > pp['avf_wt'] or count < 1 or count > pp['ag_cap'] or \
            count != len(sorted_chmsgs):
        return False
    challs = [challenge(pp, m) for m in sorted_chmsgs]
    coefs = agg_coefs(pp, agmsg, count)
    total = sum([(vk[0] * c + vk[1]) * a for a, c, vk in zip(coefs, challs, sorted_vks)])
    return pp['key_ch'] * ag_sig == total


# ------------------------------------------------------------------ adaptor signatures
def witgen_one(pp: dict, seed: str):
    """-> (witness vector, statement polynomial)"""
    wit = sample_vector(pp['secpar'], pp['lp'], pp['wit_salt'], seed, pp['wit_bd'], pp['wit_wt'])
    return wit, pp['key_ch'] * wit


def adapt(presig: PolynomialVector, wit: PolynomialVector) -> PolynomialVector:
    return presig + wit


def extract(presig: PolynomialVector, sig: PolynomialVector) -> PolynomialVector:
    return sig - presig


def witness_verify(pp: dict, wit: PolynomialVector, st: Polynomial) -> bool:
    cnws = wit.get_coef_rep()
    n, w = max(i[1] for i in cnws), max(i[2] for i in cnws)
    if n > pp['ext_wit_bd'] or w > pp['ext_wit_wt']:
        return False
    return pp['key_ch'] * wit == st
