"""GPU-backed stand-in for the ten names the reference imports from `lattice_algebra==0.1.1`
(one_time_keys.py:4-5, lm_one_time_sigs.py:3, bklm_one_time_agg_sigs.py:1, adaptor_sigs.py:1):

    LatticeParameters, Polynomial, PolynomialVector, hash2polynomial, hash2polynomialvector,
    random_polynomial, random_polynomialvector, is_bitstring, is_ntt_friendly_prime,
    UNIFORM_INFINITY_WEIGHT   (+ bits_to_indices, bits_to_decode)

Objects keep the reference's duck type (`*`, `**`, `+`, `-`, `==`, `get_coef_rep()`, `.lp`,
`.entries`, `.const_time_flag`).  Storage is an engine-format array: centred int16 coefficients
and/or the engine's uint16 NTT form; every ring operation is a call into the CUDA engine (there is
no CPU arithmetic path).  `const_time_flag` is accepted and kept but has no effect: the kernels
have no secret-dependent branches or addresses except the sampler's position select, exactly
where the reference's own decoder indexes a Python list.

`Polynomial.ntt_representation` returns the reference's raw length-2d list (2d-point cyclic transform,
parity level L3) computed on demand; it is read-only (the engine's own storage is the d-point negacyclic form).
"""
from math import ceil, isqrt, log2
from secrets import token_hex
from typing import Dict, List, Optional, Tuple

import numpy as np

from .engine import Engine

UNIFORM_INFINITY_WEIGHT: str = 'inf,wt,unif'


def coef_dtype(lp) -> type:
    """Element type of engine-format coefficient arrays: int16 for q < 2^16, int32 for wider moduli."""
    return np.int32 if lp.modulus >= 65536 else np.int16


# ------------------------------------------------------------------------------- host-side parameter logic
def bits_to_indices(secpar: int, degree: int, wt: int) -> int:
    return ceil(log2(degree)) + (wt - 1) * (ceil(log2(degree)) + secpar)


def bits_to_decode(secpar: int, bd: int) -> int:
    return ceil(log2(bd)) + 1 + secpar


def is_bitstring(val) -> bool:
    return isinstance(val, str) and ''.join(sorted(set(val))) in '01'


def is_ntt_friendly_prime(modulus: int, degree: int) -> bool:
    prime = isinstance(modulus, int) and modulus >= 2 and all(modulus % f for f in range(2, isqrt(modulus) + 1))
    pow2 = isinstance(degree, int) and degree > 0 and degree & (degree - 1) == 0
    return prime and pow2 and modulus % (2 * degree) == 1


class LatticeParameters(object):
    """LatticeParameters(modulus=, degree=, length=) as constructed at lm_one_time_sigs.py:20-21."""

    def __init__(self, degree: int, length: int, modulus: int):
        if not all(isinstance(i, int) for i in (degree, length, modulus)) or degree < 2 or length < 1:
            raise ValueError('LatticeParameters needs integer degree >= 2, length >= 1 and modulus.')
        if not is_ntt_friendly_prime(modulus=modulus, degree=degree):
            raise ValueError('LatticeParameters needs a prime modulus = 1 mod 2*degree and a power-of-two degree.')
        self.degree, self.length, self.modulus = degree, length, modulus
        self.halfmod = modulus // 2
        self.logmod = ceil(log2(modulus))
        self.n = 2 * degree
        x = 2
        while not (pow(x, 2 * degree, modulus) == 1 and pow(x, degree, modulus) != 1):
            x += 1
        self.rou, self.rou_inv = x, pow(x, 2 * degree - 1, modulus)

    def __eq__(self, other) -> bool:
        return isinstance(other, LatticeParameters) and \
            (self.degree, self.length, self.modulus) == (other.degree, other.length, other.modulus)

    def __hash__(self):
        return hash((self.degree, self.length, self.modulus))

    def __repr__(self) -> str:
        return str((self.degree, self.length, self.modulus))


# ------------------------------------------------------------------------------- engine registry
_ENGINES: Dict[Tuple[int, int, int, int, int], Engine] = {}
_DEFAULT_DEVICE = 0
_DEFAULT_SECPAR = 128


def set_default_device(device: int):
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = device


def engine_for(lp: LatticeParameters, secpar: Optional[int] = None, device: Optional[int] = None) -> Engine:
    """The (cached) CUDA context for these lattice parameters.  Ring arithmetic does not depend on
    secpar, so objects that do not know it borrow any context with the same (q, d, l)."""
    dev = _DEFAULT_DEVICE if device is None else device
    if secpar is None:
        for (sp, q, d, l, dv), eng in _ENGINES.items():
            if (q, d, l, dv) == (lp.modulus, lp.degree, lp.length, dev):
                return eng
        secpar = _DEFAULT_SECPAR
    key = (secpar, lp.modulus, lp.degree, lp.length, dev)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(secpar, lp.modulus, lp.degree, lp.length, device=dev)
    return _ENGINES[key]


def ensure_key_ch(eng: Engine, key_ch: 'PolynomialVector'):
    """Make `key_ch` the NTT-resident public row of the context (uploaded once per distinct row)."""
    eng.ensure_key_ch(key_ch.coef)


class BoundEngine(object):
    """An Engine paired with one public row: every method call takes the engine's lock, makes `key_ch` the resident
    row if another pp (or another thread) replaced it, and then runs - so two parameter sets that share a cached
    context, or two threads, cannot hand each other's key_ch to a keygen / verify call."""

    def __init__(self, eng: Engine, key_ch_coef: np.ndarray):
        self._eng, self._coef = eng, key_ch_coef

    def __getattr__(self, name):
        attr = getattr(self._eng, name)
        if not callable(attr):
            return attr
        eng, coef = self._eng, self._coef

        def call(*a, **k):
            with eng.lock:
                eng.ensure_key_ch(coef)
                return attr(*a, **k)
        return call


# ------------------------------------------------------------------------------- ring elements
class Polynomial(object):
    """Element of Z_q[X]/(X^d+1).  `coefs` as in lattice_algebra: {index: centred coefficient}."""

    def __init__(self, lp: LatticeParameters, coefs: Optional[Dict[int, int]] = None, const_time_flag: bool = True,
                 _coef: Optional[np.ndarray] = None, _ntt: Optional[np.ndarray] = None):
        if not isinstance(lp, LatticeParameters):
            raise ValueError('Polynomial needs LatticeParameters.')
        self.lp = lp
        self.const_time_flag = const_time_flag
        self._coef, self._ntt = _coef, _ntt
        if _coef is None and _ntt is None:
            if not isinstance(coefs, dict):
                raise ValueError('Polynomial needs a coefficient dictionary.')
            dense = np.zeros(lp.degree, dtype=coef_dtype(lp))
            for i, v in coefs.items():
                if not isinstance(i, int) or not isinstance(v, int) or not 0 <= i < lp.degree or abs(v) > lp.halfmod:
                    raise ValueError('Polynomial coefficient index or magnitude out of range.')
                dense[i] = v
            self._coef = dense

    # engine-format views (converted on the GPU on first use)
    @property
    def coef(self) -> np.ndarray:
        if self._coef is None:
            self._coef = engine_for(self.lp).ntt_inv(np.ascontiguousarray(self._ntt[None]))[0]
        return self._coef

    @property
    def ntt(self) -> np.ndarray:
        if self._ntt is None:
            self._ntt = engine_for(self.lp).ntt_fwd(np.ascontiguousarray(self._coef[None]))[0]
        return self._ntt

    @property
    def ntt_representation(self) -> List[int]:
        """The reference's own storage (parity level L3): 2d-point cyclic transform of the zero-padded
        coefficient list, natural order, centred residues - computed on the GPU on demand."""
        return engine_for(self.lp).ntt_reference_repr(np.ascontiguousarray(self.coef[None]))[0].tolist()

    def get_coef_rep(self) -> Tuple[Dict[int, int], int, int]:
        c = self.coef
        nz = np.flatnonzero(c)
        if nz.size == 0:
            return {}, 0, 0
        return {int(i): int(c[i]) for i in nz}, int(np.abs(c.astype(np.int64)).max()), int(nz.size)

    def __eq__(self, other) -> bool:
        return isinstance(other, Polynomial) and self.lp == other.lp and np.array_equal(self.coef, other.coef)

    def __bool__(self) -> bool:
        return True

    def _wrap(self, coef: np.ndarray) -> 'Polynomial':
        return Polynomial(self.lp, const_time_flag=self.const_time_flag, _coef=coef)

    def __add__(self, other):
        if isinstance(other, int) and other == 0:
            return self
        return self._wrap(engine_for(self.lp).vec_add(self.coef[None], other.coef[None])[0])

    __radd__ = __add__

    def __sub__(self, other):
        return self._wrap(engine_for(self.lp).vec_sub(self.coef[None], other.coef[None])[0])

    def __neg__(self):
        zero = np.zeros((1, self.lp.degree), dtype=coef_dtype(self.lp))
        return self._wrap(engine_for(self.lp).vec_sub(zero, self.coef[None])[0])

    def __mul__(self, other):
        if isinstance(other, int) and other == 0:
            return 0
        return self._wrap(engine_for(self.lp).poly_mul(self.coef[None], other.coef[None])[0])

    __rmul__ = __mul__

    def __repr__(self) -> str:
        return str(self.get_coef_rep())


class PolynomialVector(object):
    """`*` is the dot product (-> Polynomial), `**` scales every entry by a Polynomial."""

    def __init__(self, lp: LatticeParameters, entries: Optional[List[Polynomial]] = None, const_time_flag: bool = True,
                 _coef: Optional[np.ndarray] = None, _ntt: Optional[np.ndarray] = None):
        if not isinstance(lp, LatticeParameters):
            raise ValueError('PolynomialVector needs LatticeParameters.')
        self.lp = lp
        self.const_time_flag = const_time_flag
        self._entries = None
        self._coef, self._ntt = _coef, _ntt
        if _coef is None and _ntt is None:
            if not isinstance(entries, list) or not all(isinstance(i, Polynomial) and i.lp == lp for i in entries):
                raise ValueError('PolynomialVector needs a list of Polynomials over the same LatticeParameters.')
            self._entries = entries
            self._coef = np.stack([e.coef for e in entries]) if entries else np.zeros((0, lp.degree), coef_dtype(lp))

    @property
    def coef(self) -> np.ndarray:
        if self._coef is None:
            self._coef = engine_for(self.lp).ntt_inv(np.ascontiguousarray(self._ntt))
        return self._coef

    @property
    def ntt(self) -> np.ndarray:
        if self._ntt is None:
            self._ntt = engine_for(self.lp).ntt_fwd(np.ascontiguousarray(self._coef))
        return self._ntt

    @property
    def entries(self) -> List[Polynomial]:
        if self._entries is None:
            c = self.coef
            self._entries = [Polynomial(self.lp, const_time_flag=self.const_time_flag, _coef=c[i]) for i in range(c.shape[0])]
        return self._entries

    def __len__(self):
        return int((self._coef if self._coef is not None else self._ntt).shape[0])

    def get_coef_rep(self) -> List[Tuple[Dict[int, int], int, int]]:
        return [e.get_coef_rep() for e in self.entries]

    def __eq__(self, other) -> bool:
        return isinstance(other, PolynomialVector) and self.lp == other.lp and np.array_equal(self.coef, other.coef)

    def __bool__(self) -> bool:
        return True

    def _wrap(self, coef: np.ndarray) -> 'PolynomialVector':
        return PolynomialVector(self.lp, const_time_flag=self.const_time_flag, _coef=coef)

    def __add__(self, other):
        if isinstance(other, int) and other == 0:
            return self
        return self._wrap(engine_for(self.lp).vec_add(self.coef, other.coef))

    __radd__ = __add__

    def __sub__(self, other):
        return self._wrap(engine_for(self.lp).vec_sub(self.coef, other.coef))

    def __mul__(self, other) -> Polynomial:
        eng = engine_for(self.lp)
        prods = eng.poly_mul(self.coef, other.coef)                      # entry-wise products
        while prods.shape[0] > 1:                                        # tree sum on the GPU
            half = prods.shape[0] // 2
            head = eng.vec_add(np.ascontiguousarray(prods[:half]), np.ascontiguousarray(prods[half:2 * half]))
            prods = np.concatenate([head, prods[2 * half:]]) if prods.shape[0] % 2 else head
        return Polynomial(self.lp, const_time_flag=self.const_time_flag, _coef=prods[0])

    def __pow__(self, scalar: Polynomial):
        rep = np.ascontiguousarray(np.broadcast_to(scalar.coef, self.coef.shape))
        return self._wrap(engine_for(self.lp).poly_mul(self.coef, rep))

    def __repr__(self) -> str:
        return str(self.entries)


# ------------------------------------------------------------------------------- samplers
def _check(distribution: str, dist_pars: Dict[str, int], num_coefs: int, lp: LatticeParameters):
    if distribution != UNIFORM_INFINITY_WEIGHT:
        raise ValueError('Unsupported distribution.')
    bd, wt = dist_pars['bd'], dist_pars['wt']
    if not 1 <= bd <= lp.halfmod or not 1 <= wt <= lp.degree or num_coefs != wt:
        raise ValueError('Cannot sample with these bound / weight parameters.')
    return bd, wt


def hash2polynomialvector(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                          num_coefs: int, bti: int, btd: int, msg: str, salt: str,
                          const_time_flag: bool = True) -> PolynomialVector:
    bd, wt = _check(distribution, dist_pars, num_coefs, lp)
    dense, _ = engine_for(lp, secpar).hash2polyvec(salt, [msg], bd, wt, lp.length)
    return PolynomialVector(lp, const_time_flag=const_time_flag, _coef=dense[0])


def hash2polynomial(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int], salt: str,
                    msg: str, num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> Polynomial:
    bd, wt = _check(distribution, dist_pars, num_coefs, lp)
    dense, _ = engine_for(lp, secpar).hash2polyvec(salt, [msg], bd, wt, 1)
    return Polynomial(lp, const_time_flag=const_time_flag, _coef=dense[0, 0])


def random_polynomialvector(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                            num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> PolynomialVector:
    """The reference feeds its decoder `secrets.randbits`; here a fresh 256-bit secret from `secrets`
    is expanded by SHAKE256 on the GPU into the same decoder (same distribution, not reproducible,
    exactly like the reference's)."""
    return hash2polynomialvector(secpar, lp, distribution, dist_pars, num_coefs, bti, btd, msg=token_hex(32),
                                 salt='RANDOM_POLYNOMIALVECTOR', const_time_flag=const_time_flag)


def random_polynomial(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                      num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> Polynomial:
    return hash2polynomial(secpar, lp, distribution, dist_pars, salt='RANDOM_POLYNOMIAL', msg=token_hex(32),
                           num_coefs=num_coefs, bti=bti, btd=btd, const_time_flag=const_time_flag)
