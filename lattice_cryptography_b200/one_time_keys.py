"""Key / parameter containers with the reference's names, attributes, validation order and error
strings (reference lattice_cryptography/one_time_keys.py:20-299), over the GPU-backed ring objects
of `lattice_cryptography_b200.lattice_algebra`.  Host-only logic: no arithmetic happens here except
sampling `key_ch` (SchemeParameters), which runs the sampler kernel.

The reference defines the bit-budget helpers as `bits_per_index_set` / `bits_per_coefficient`
(:243-256) while its other modules import them as `bits_to_indices` / `bits_to_decode`; both
spellings are exported here.
"""
from math import ceil, log2

from .lattice_algebra import (LatticeParameters, Polynomial, PolynomialVector, UNIFORM_INFINITY_WEIGHT,
                              is_bitstring, random_polynomialvector)

ALLOWABLE_SECPARS = [128, 256]
ALLOWABLE_DISTRIBUTIONS = [UNIFORM_INFINITY_WEIGHT]

# error strings (one_time_keys.py:12-17 and the per-class constants)
GENERIC_ERR = 'Something went wrong.'
MISSING_DATA_ERR = 'Missing some required data.'
INCORRECT_DATA_TYPE_ERR = 'Required input data not the correct type.'
DATA_MISMATCH_ERR = 'Input data did not match.'
SEED_INST_ERR_NEED_BITS = INCORRECT_DATA_TYPE_ERR + ' Input must be a binary string.'
INVALID_DATA_VALUES_ERR = 'Required input data does not have valid values.'
SECWIT_INST_ERR_LP_MISMATCH = DATA_MISMATCH_ERR + (' Input LatticeParameters object does not match the '
                                                   'LatticeParameters for the input PolynomialVector.')
SK_KEY_LP_MISMATCH = DATA_MISMATCH_ERR + ' Input LatticeParameters objects or security parameter integers do not match.'
VK_LP_OR_SECPAR_MISMATCH = SK_KEY_LP_MISMATCH


def _bad_secpar(secpar, wording: str) -> bool:
    """True (after raising is left to the caller) when secpar is not an allowed integer."""
    return not isinstance(secpar, int) or secpar not in ALLOWABLE_SECPARS


def _secpar_msg(secpar, middle: str) -> str:
    return INVALID_DATA_VALUES_ERR + f' Input security parameter must be{middle} {ALLOWABLE_SECPARS} but had {secpar}.'


# Opt-in: when True, str() of a verification key / public statement is a digest of its CONTENT
# instead of CPython's address-based default (reference one_time_keys.py:197-237 defines no __str__).  The string
# is part of every challenge hash input (lm_one_time_sigs.py:148, adaptor_sigs.py:176,
# bklm_one_time_agg_sigs.py:65), so switching it on changes the signatures - and makes them verifiable against a
# pickled / re-created key object, which the reference's signatures are not.  Off by default.
CONTENT_STR: bool = False


def set_content_str(on: bool) -> None:
    global CONTENT_STR
    CONTENT_STR = bool(on)


def _content_digest(name: str, secpar: int, polys) -> str:
    from hashlib import shake_256
    h = shake_256(name.encode() + secpar.to_bytes(2, 'little'))
    for f in polys:
        h.update(f.coef.astype('<i2').tobytes())       # centred coefficients, natural order, little-endian int16
    return f'<{name} {h.hexdigest(16)}>'


class _Comparable(object):
    """__eq__/__bool__ over a fixed tuple of attribute names (the reference compares attribute by attribute)."""
    _fields = ()

    def __eq__(self, other) -> bool:
        return all(getattr(self, f) == getattr(other, f) for f in self._fields)

    def __bool__(self) -> bool:
        return all(bool(getattr(self, f)) for f in self._fields)

    __hash__ = object.__hash__


class SecretSeed(_Comparable):
    _fields = ('secpar', 'lp', 'seed')

    def __init__(self, seed: str, secpar: int, lp: LatticeParameters):
        if _bad_secpar(secpar, ''):
            raise ValueError(_secpar_msg(secpar, ' an integer in'))
        elif not is_bitstring(seed):
            raise ValueError(SEED_INST_ERR_NEED_BITS)
        elif not isinstance(lp, LatticeParameters):
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Input lattice parameters must be LatticeParameters object.')
        elif len(seed) < secpar:
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Input secret seed must have enough bits.')
        self.secpar, self.lp, self.seed = secpar, lp, seed


class OneTimeSecretWitness(_Comparable):
    _fields = ('secpar', 'lp', 'key')

    def __init__(self, secpar: int, lp: LatticeParameters, key: PolynomialVector):
        if secpar not in ALLOWABLE_SECPARS:
            raise ValueError(_secpar_msg(secpar, ' in'))
        elif key.lp != lp:
            raise ValueError(SECWIT_INST_ERR_LP_MISMATCH)
        self.secpar, self.lp, self.key = secpar, lp, key
        self.key.const_time_flag = True          # secret material: reference one_time_keys.py:82-83


class OneTimePublicStatement(_Comparable):
    _fields = ('secpar', 'lp', 'key')

    def __init__(self, secpar: int, lp: LatticeParameters, key: Polynomial):
        if _bad_secpar(secpar, ''):
            raise ValueError(_secpar_msg(secpar, ' in'))
        elif not isinstance(lp, LatticeParameters):
            raise ValueError(INVALID_DATA_VALUES_ERR + f' Input lattice parameters must be LatticeParameters but had {type(lp)}.')
        elif not isinstance(key, Polynomial):
            raise ValueError(INVALID_DATA_VALUES_ERR + f' Input key must be Polynomial but had {type(key)}.')
        elif key.lp != lp:
            raise ValueError(SECWIT_INST_ERR_LP_MISMATCH)
        self.secpar, self.lp, self.key = secpar, lp, key
        self.key.const_time_flag = False         # public material: reference one_time_keys.py:126

    def __repr__(self) -> str:          # str() falls back to this; repr() is what str(list(...)) uses (BKLM)
        return _content_digest('OneTimePublicStatement', self.secpar, [self.key]) if CONTENT_STR else object.__repr__(self)


class _LeftRight(_Comparable):
    """A (left_key, right_key) pair addressed as [0] / [1]."""
    _fields = ('secpar', 'lp', 'left_key', 'right_key')
    _flag = True

    def _install(self, secpar, lp, left_key, right_key):
        self.secpar, self.lp, self.left_key, self.right_key = secpar, lp, left_key, right_key
        self.left_key.const_time_flag = self._flag
        self.right_key.const_time_flag = self._flag

    def __getitem__(self, item: int):
        if item not in [0, 1]:
            raise ValueError('Can only get two items.')
        return self.right_key if item else self.left_key


class OneTimeSigningKey(_LeftRight):
    _flag = True

    def __init__(self, secpar: int, lp: LatticeParameters, left_key: PolynomialVector, right_key: PolynomialVector):
        if _bad_secpar(secpar, ''):
            raise ValueError(_secpar_msg(secpar, ' in'))
        elif not isinstance(lp, LatticeParameters):
            raise ValueError(INVALID_DATA_VALUES_ERR + f' Input lattice parameters must be a LatticeParameters object, but had {type(lp)}.')
        elif not isinstance(left_key, PolynomialVector) or not isinstance(right_key, PolynomialVector):
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Both input keys must be PolynomialVectors.')
        elif left_key.lp != lp or right_key.lp != lp:
            raise ValueError(SECWIT_INST_ERR_LP_MISMATCH)
        self._install(secpar, lp, left_key, right_key)


class OneTimeVerificationKey(_LeftRight):
    _flag = False

    def __init__(self, secpar: int, lp: LatticeParameters, left_key: Polynomial, right_key: Polynomial):
        if _bad_secpar(secpar, ''):
            raise ValueError(_secpar_msg(secpar, ' in'))
        elif not isinstance(lp, LatticeParameters):
            raise ValueError(INVALID_DATA_VALUES_ERR + f' Input lattice parameters must be LatticeParameters, but had {type(lp)}.')
        elif not isinstance(left_key, Polynomial) or not isinstance(right_key, Polynomial):
            raise ValueError(INVALID_DATA_VALUES_ERR + f' Both input keys must be Polynomial but had {type(left_key)} and {type(right_key)}.')
        elif left_key.lp != lp or right_key.lp != lp:
            raise ValueError(SECWIT_INST_ERR_LP_MISMATCH)
        self._install(secpar, lp, left_key, right_key)

    def __repr__(self) -> str:          # str() falls back to this; repr() is what str(list(...)) uses (BKLM)
        if CONTENT_STR:
            return _content_digest('OneTimeVerificationKey', self.secpar, [self.left_key, self.right_key])
        return object.__repr__(self)


def bits_per_index_set(secpar: int, degree: int, wt: int) -> int:
    """Bits needed to draw wt distinct positions out of `degree` with bias O(2**-secpar)."""
    return ceil(log2(degree)) + (wt - 1) * (ceil(log2(degree)) + secpar)


def bits_per_coefficient(secpar: int, bd: int) -> int:
    """Bits needed to draw a coefficient from [-bd, bd] with bias O(2**-secpar)."""
    if bd <= 0:
        raise ValueError('Cannot compute bits per coefficient for a non-positive bound bd.')
    return ceil(log2(bd)) + 1 + secpar


bits_to_indices = bits_per_index_set       # the names the reference's other modules import
bits_to_decode = bits_per_coefficient


class SchemeParameters(object):
    def __init__(self, secpar: int, lp: LatticeParameters, distribution: str, key_ch: PolynomialVector = None):
        if _bad_secpar(secpar, ''):
            raise ValueError(_secpar_msg(secpar, ' in'))
        elif not isinstance(lp, LatticeParameters):
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Input lattice parameters must be LatticeParameters.')
        elif key_ch is not None and not isinstance(key_ch, PolynomialVector):
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Input key challenge must be a PolynomialVector or None.')
        elif not isinstance(distribution, str) or distribution not in ALLOWABLE_DISTRIBUTIONS:
            raise ValueError(INVALID_DATA_VALUES_ERR + ' Input distribution must be a string code indicating a supported distribution.')
        elif key_ch is not None and key_ch.lp != lp:
            raise ValueError(SECWIT_INST_ERR_LP_MISMATCH)
        self.secpar, self.lp, self.distribution = secpar, lp, distribution
        if key_ch is not None:
            self.key_ch = key_ch
            self.key_ch.const_time_flag = False
        else:
            # uniform public row: every position non-zero, coefficients in +-[1 .. q//2]
            self.key_ch = random_polynomialvector(
                secpar=secpar, lp=lp, distribution=distribution, dist_pars={'bd': lp.modulus // 2, 'wt': lp.degree},
                bti=bits_per_index_set(secpar=secpar, degree=lp.degree, wt=lp.degree),
                btd=bits_per_coefficient(secpar=secpar, bd=lp.modulus // 2), const_time_flag=True,
                num_coefs=lp.degree)

    def __eq__(self, other) -> bool:
        return self.secpar == other.secpar and self.lp == other.lp and self.key_ch == other.key_ch and \
            self.distribution == other.distribution

    __hash__ = object.__hash__
