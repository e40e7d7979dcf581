"""CPU ORACLE (test infrastructure, never the product path): a restatement of the
third-party package ``lattice_algebra==0.1.1`` (PyPI ``lattice-algebra``), which the
reference pins at requirements.txt:1 / setup.py:44 and imports at
lattice_cryptography/one_time_keys.py:4-5, lm_one_time_sigs.py:3,
bklm_one_time_agg_sigs.py:1 and adaptor_sigs.py:1.

PARITY UNPINNED: that package is not vendored under /root/reference, is not installed
here and cannot be downloaded (no network), and the reference ships no golden vectors or
known-answer tests for it.  This file restates its published
algorithm from the reference's call sites and from the semantic items U1-U14 listed in
SURVEY.md section 8(c).  Every U-item lives in exactly one named function below so that a
maintainer holding the real package can diff them one by one.  SHAKE256 itself is pinned
by ``hashlib``.

It exists so that the reference's own scheme modules (imported unmodified from
/root/reference by oracle/ref_loader.py) and the restated scheme modules in
oracle/schemes.py can run on a CPU and serve as the bit-exact checker for the CUDA
engine.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it.

Deliberately written for obviousness, with Python integers and lists, mirroring the cost
structure of the original (2d-point cyclic transform, centred residues).
"""
from copy import deepcopy
from hashlib import shake_256
from math import ceil, log2, isqrt
from secrets import randbits
from typing import Dict, List, Tuple, Union

UNIFORM_INFINITY_WEIGHT: str = 'inf,wt,unif'

__all__ = [
    'UNIFORM_INFINITY_WEIGHT', 'LatticeParameters', 'Polynomial', 'PolynomialVector',
    'hash2polynomial', 'hash2polynomialvector', 'random_polynomial', 'random_polynomialvector',
    'is_bitstring', 'is_ntt_friendly_prime', 'bits_to_indices', 'bits_to_decode',
    'binary_digest', 'decode2indices', 'decode2coef', 'decode2polycoefs', 'get_gen_bytes_per_poly',
    'cent', 'ntt',
]


# --------------------------------------------------------------------------------------
# bit budgets (same formulas as reference one_time_keys.py:243-256)
# --------------------------------------------------------------------------------------
def bits_to_indices(secpar: int, degree: int, wt: int) -> int:
    return ceil(log2(degree)) + (wt - 1) * (ceil(log2(degree)) + secpar)


def bits_to_decode(secpar: int, bd: int) -> int:
    return ceil(log2(bd)) + 1 + secpar


# --------------------------------------------------------------------------------------
# number theory helpers  (U7, U12)
# --------------------------------------------------------------------------------------
def is_prime(val: int) -> bool:
    if val < 2:
        return False
    return all(val % f for f in range(2, isqrt(val) + 1))


def is_pow_two(val: int) -> bool:
    return isinstance(val, int) and val > 0 and val & (val - 1) == 0


def has_prim_rou(modulus: int, degree: int) -> bool:
    return modulus % (2 * degree) == 1


def is_ntt_friendly_prime(modulus: int, degree: int) -> bool:
    return is_prime(modulus) and is_pow_two(degree) and has_prim_rou(modulus=modulus, degree=degree)


def is_prim_rou(modulus: int, degree: int, val: int) -> bool:
    """val has multiplicative order exactly 2*degree (a power of two, so it suffices to
    look at the half-order power)."""
    return pow(val, 2 * degree, modulus) == 1 and pow(val, degree, modulus) != 1


def get_prim_rou_and_rou_inv(modulus: int, degree: int) -> Tuple[int, int]:
    """U7: the least x >= 2 of order exactly 2*degree, and its inverse x**(2d-1)."""
    if not is_ntt_friendly_prime(modulus=modulus, degree=degree):
        raise ValueError('Input modulus and degree must be ntt-friendly.')
    x = 2
    while x < modulus and not is_prim_rou(modulus, degree, x):
        x += 1
    return x, pow(x, 2 * degree - 1, modulus)


def is_bitstring(val) -> bool:
    """U12: the empty string counts as a bitstring."""
    return isinstance(val, str) and ''.join(sorted(set(val))) in '01'


def bit_rev(num_bits: int, val: int) -> int:
    return int(bin(val)[2:].zfill(num_bits)[::-1], 2)


def bit_rev_cp(val: List[int], num_bits: int) -> List[int]:
    return [val[bit_rev(num_bits, i)] for i in range(len(val))]


def cent(q: int, halfmod: int, logmod: int, val: int) -> int:
    """U7: constant-time centring of val mod q into [-(q-1)/2, (q-1)/2] (odd q)."""
    y = val % q
    return y - (1 + ((y - halfmod - 1) >> logmod)) * q


def _cent_branchy(q: int, halfmod: int, val: int) -> int:
    """U14: the non-constant-time path; numerically identical to cent()."""
    y = val % q
    return y - q if y > halfmod else y


def make_zetas_and_invs(q: int, d: int, halfmod: int, logmod: int, n: int, lgn: int) -> Tuple[List[int], List[int]]:
    """U7: stage s (block size m = 2**s) uses zeta**(n/m)."""
    zeta, zeta_inv = get_prim_rou_and_rou_inv(modulus=q, degree=d)
    exps = [n >> (s + 1) for s in range(lgn)]
    return ([cent(q, halfmod, logmod, pow(zeta, e, q)) for e in exps],
            [cent(q, halfmod, logmod, pow(zeta_inv, e, q)) for e in exps])


def ntt(q: int, zetas: List[int], zetas_inv: List[int], inv_flag: bool, halfmod: int, logmod: int, n: int,
        lgn: int, val: List[int], const_time_flag: bool = True) -> List[int]:
    """U7: n-point *cyclic* Cooley-Tukey transform (n = 2d): bit-reverse the input, then
    lgn stages of butterflies with a running twiddle; natural-order output; every stored
    intermediate is centred.  The inverse uses zeta**-1 and a final multiplication by
    n**-1.  (The running twiddle is reduced here; the original lets it grow as a Python
    big integer, which changes nothing numerically because every use is reduced.)"""
    if len(val) != n:
        raise ValueError('Can only transform vectors of length n.')
    x = bit_rev_cp(val, lgn)
    m = 1
    for s in range(1, lgn + 1):
        m *= 2
        this_zeta = zetas_inv[s - 1] if inv_flag else zetas[s - 1]
        half = m // 2
        for k in range(0, n, m):
            w = 1
            for j in range(half):
                t = w * x[k + j + half]
                u = x[k + j]
                if const_time_flag:
                    x[k + j] = cent(q, halfmod, logmod, u + t)
                    x[k + j + half] = cent(q, halfmod, logmod, u - t)
                else:
                    x[k + j] = _cent_branchy(q, halfmod, u + t)
                    x[k + j + half] = _cent_branchy(q, halfmod, u - t)
                w = (w * this_zeta) % q
    if inv_flag:
        n_inv = pow(n, q - 2, q)
        x = [cent(q, halfmod, logmod, n_inv * i) for i in x]
    return x


def binary_digest(msg: str, num_bytes: int, salt: str) -> str:
    """U1 + U2: SHAKE256(salt || msg), salt first, UTF-8; digest read as one big-endian
    integer and printed MSB-first, zero-filled to 8*num_bytes characters."""
    m = shake_256()
    m.update(salt.encode() + msg.encode())
    return bin(int(m.hexdigest(num_bytes), 16))[2:].zfill(8 * num_bytes)


# --------------------------------------------------------------------------------------
# parameters and ring elements
# --------------------------------------------------------------------------------------
class LatticeParameters(object):
    """U7.  Keyword signature pinned by reference tests/test_one_time_keys.py:31."""
    degree: int
    length: int
    modulus: int

    def __init__(self, degree: int, length: int, modulus: int):
        if not isinstance(degree, int) or not isinstance(length, int) or not isinstance(modulus, int):
            raise ValueError('LatticeParameters needs integer degree, length and modulus.')
        if degree < 2 or length < 1 or modulus < 3:
            raise ValueError('LatticeParameters needs degree >= 2, length >= 1, modulus >= 3.')
        if not is_ntt_friendly_prime(modulus=modulus, degree=degree):
            raise ValueError('LatticeParameters needs a prime modulus = 1 mod 2*degree and a power-of-two degree.')
        self.degree = degree
        self.length = length
        self.modulus = modulus
        self.halfmod = modulus // 2
        self.logmod = ceil(log2(modulus))
        self.n = 2 * degree
        self.lgn = ceil(log2(self.n))
        self.rou, self.rou_inv = get_prim_rou_and_rou_inv(modulus=modulus, degree=degree)
        self.zetas, self.zetas_invs = make_zetas_and_invs(modulus, degree, self.halfmod, self.logmod, self.n,
                                                          self.lgn)

    def __eq__(self, other) -> bool:
        return isinstance(other, LatticeParameters) and \
            (self.degree, self.length, self.modulus) == (other.degree, other.length, other.modulus)

    def __hash__(self):
        return hash((self.degree, self.length, self.modulus))

    def __repr__(self) -> str:
        return str((self.degree, self.length, self.modulus))


class Polynomial(object):
    """Element of Z_q[X]/(X^d+1) held as the 2d-point cyclic transform of its zero-padded
    coefficient list (U7, U13)."""
    lp: LatticeParameters
    ntt_representation: List[int]
    const_time_flag: bool

    def __init__(self, lp: LatticeParameters, coefs: Dict[int, int], const_time_flag: bool = True):
        if not isinstance(lp, LatticeParameters) or not isinstance(coefs, dict):
            raise ValueError('Polynomial needs LatticeParameters and a coefficient dictionary.')
        for i, v in coefs.items():
            if not isinstance(i, int) or not isinstance(v, int) or not 0 <= i < lp.degree or abs(v) > lp.halfmod:
                raise ValueError('Polynomial coefficient index or magnitude out of range.')
        self.lp = lp
        self.const_time_flag = const_time_flag
        padded = [0] * lp.n
        for i, v in coefs.items():
            padded[i] = v
        self.ntt_representation = self._ntt(inv_flag=False, val=padded)

    def _ntt(self, inv_flag: bool, val: List[int]) -> List[int]:
        lp = self.lp
        return ntt(q=lp.modulus, zetas=lp.zetas, zetas_inv=lp.zetas_invs, inv_flag=inv_flag, halfmod=lp.halfmod,
                   logmod=lp.logmod, n=lp.n, lgn=lp.lgn, val=val, const_time_flag=self.const_time_flag)

    def _c(self, v: int) -> int:
        lp = self.lp
        return cent(lp.modulus, lp.halfmod, lp.logmod, v)

    def get_coef_rep(self) -> Tuple[Dict[int, int], int, int]:
        """U9: inverse transform, fold lower - upper, centre; ({i: v != 0}, inf-norm, weight)."""
        d = self.lp.degree
        tmp = self._ntt(inv_flag=True, val=self.ntt_representation)
        folded = [self._c(lo - hi) for lo, hi in zip(tmp[:d], tmp[d:])]
        coefs = {i: v for i, v in enumerate(folded) if v != 0}
        if not coefs:
            return coefs, 0, 0
        return coefs, max(abs(v) for v in coefs.values()), len(coefs)

    def __eq__(self, other) -> bool:
        """U8: equality is on (lp, coefficient representation)."""
        if not isinstance(other, Polynomial) or self.lp != other.lp:
            return False
        return self.get_coef_rep() == other.get_coef_rep()

    def __bool__(self) -> bool:
        return True

    def __add__(self, other):
        if isinstance(other, int) and other == 0:  # U11: sum() starts from 0
            return self
        result = deepcopy(self)
        result.ntt_representation = [self._c(x + y) for x, y in
                                     zip(self.ntt_representation, other.ntt_representation)]
        return result

    def __radd__(self, other):
        return self.__add__(other)

    def __sub__(self, other):
        result = deepcopy(self)
        result.ntt_representation = [self._c(x - y) for x, y in
                                     zip(self.ntt_representation, other.ntt_representation)]
        return result

    def __neg__(self):
        result = deepcopy(self)
        result.ntt_representation = [-x for x in self.ntt_representation]
        return result

    def __mul__(self, other):
        if isinstance(other, int) and other == 0:  # U11
            return 0
        result = deepcopy(self)
        result.ntt_representation = [self._c(x * y) for x, y in
                                     zip(self.ntt_representation, other.ntt_representation)]
        return result

    def __rmul__(self, other):
        return self.__mul__(other)

    def __repr__(self) -> str:
        return str(self.get_coef_rep())


class PolynomialVector(object):
    """U11: ``*`` is the dot product (-> Polynomial), ``**`` scales by a Polynomial."""
    lp: LatticeParameters
    entries: List[Polynomial]
    const_time_flag: bool

    def __init__(self, lp: LatticeParameters, entries: List[Polynomial], const_time_flag: bool = True):
        if not isinstance(lp, LatticeParameters) or not isinstance(entries, list) or \
                not all(isinstance(i, Polynomial) and i.lp == lp for i in entries):
            raise ValueError('PolynomialVector needs LatticeParameters and a list of Polynomials over them.')
        self.lp = lp
        self.entries = entries
        self.const_time_flag = const_time_flag

    def __eq__(self, other) -> bool:
        return isinstance(other, PolynomialVector) and self.lp == other.lp and \
            len(self.entries) == len(other.entries) and all(x == y for x, y in zip(self.entries, other.entries))

    def __bool__(self) -> bool:
        return True

    def __add__(self, other):
        if isinstance(other, int) and other == 0:
            return self
        result = deepcopy(self)
        result.entries = [x + y for x, y in zip(self.entries, other.entries)]
        return result

    def __radd__(self, other):
        return self.__add__(other)

    def __sub__(self, other):
        result = deepcopy(self)
        result.entries = [x - y for x, y in zip(self.entries, other.entries)]
        return result

    def __mul__(self, other) -> Polynomial:
        return sum(x * y for x, y in zip(self.entries, other.entries))

    def __pow__(self, scalar: Polynomial):
        result = deepcopy(self)
        result.entries = [scalar * i for i in result.entries]
        return result

    def get_coef_rep(self) -> List[Tuple[Dict[int, int], int, int]]:
        return [i.get_coef_rep() for i in self.entries]

    def __repr__(self) -> str:
        return str(self.entries)


# --------------------------------------------------------------------------------------
# bitstring -> (index set, coefficients) decoder  (U3-U6)
# --------------------------------------------------------------------------------------
def decode2coef(secpar: int, lp: LatticeParameters, val: str, bd: int, btd: int) -> int:
    """U6: first bit is the sign; magnitude 1 + (remaining bits mod bd); bd == 1 -> +-1."""
    if bd < 1 or bd > lp.halfmod or len(val) < btd:
        raise ValueError('Cannot decode a coefficient with these inputs.')
    sign = 2 * int(val[0]) - 1
    if bd == 1:
        return sign
    return sign * (1 + int(val[1:btd], 2) % bd)


def decode2indices(secpar: int, lp: LatticeParameters, num_coefs: int, val: str, bti: int) -> List[int]:
    """U5: first index = the first ceil(log2 d) bits as they are; each later index picks
    remaining[r mod len(remaining)] from the ascending list of unused positions, r being
    the next ceil(log2 d)+secpar bits."""
    k = ceil(log2(lp.degree))
    if num_coefs < 1 or num_coefs > lp.degree or len(val) < bti:
        raise ValueError('Cannot decode an index set with these inputs.')
    remaining = list(range(lp.degree))
    result = [remaining.pop(int(val[:k], 2))]
    step = k + secpar
    for i in range(num_coefs - 1):
        r = int(val[k + i * step: k + (i + 1) * step], 2)
        result.append(remaining.pop(r % len(remaining)))
    return result


def decode2polycoefs(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int], val: str,
                     num_coefs: int, bti: int, btd: int) -> Dict[int, int]:
    """U4: bits [0, bti) -> indices, bits [bti, bti + num_coefs*btd) -> coefficients, paired in
    draw order; trailing pad bits ignored."""
    if distribution != UNIFORM_INFINITY_WEIGHT:
        raise ValueError('Unsupported distribution.')
    if len(val) < bti + num_coefs * btd:
        raise ValueError('Not enough bits to decode a polynomial.')
    indices = decode2indices(secpar, lp, num_coefs, val[:bti], bti)
    body = val[bti:]
    coefs = [decode2coef(secpar, lp, body[i * btd:(i + 1) * btd], dist_pars['bd'], btd) for i in range(num_coefs)]
    return {i: c for i, c in zip(indices, coefs)}


def get_gen_bytes_per_poly(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                           num_coefs: int, bti: int, btd: int) -> int:
    if distribution != UNIFORM_INFINITY_WEIGHT:
        raise ValueError('Unsupported distribution.')
    return ceil((bti + num_coefs * btd) / 8)


def hash2polynomial(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int], salt: str,
                    msg: str, num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> Polynomial:
    nb = get_gen_bytes_per_poly(secpar, lp, distribution, dist_pars, num_coefs, bti, btd)
    bits = binary_digest(msg, nb, salt)
    return Polynomial(lp=lp, coefs=decode2polycoefs(secpar, lp, distribution, dist_pars, bits, num_coefs, bti, btd),
                      const_time_flag=const_time_flag)


def hash2polynomialvector(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                          num_coefs: int, bti: int, btd: int, msg: str, salt: str,
                          const_time_flag: bool = True) -> PolynomialVector:
    """U3: ONE digest of length*nb bytes cut into `length` consecutive 8*nb-bit chunks."""
    nb = get_gen_bytes_per_poly(secpar, lp, distribution, dist_pars, num_coefs, bti, btd)
    bits = binary_digest(msg, nb * lp.length, salt)
    entries = [Polynomial(lp=lp, coefs=decode2polycoefs(secpar, lp, distribution, dist_pars,
                                                        bits[i * 8 * nb:(i + 1) * 8 * nb], num_coefs, bti, btd),
                          const_time_flag=const_time_flag) for i in range(lp.length)]
    return PolynomialVector(lp=lp, entries=entries, const_time_flag=const_time_flag)


def random_polynomial(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                      num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> Polynomial:
    """U12: the same decoder fed by secrets.randbits."""
    nbits = 8 * get_gen_bytes_per_poly(secpar, lp, distribution, dist_pars, num_coefs, bti, btd)
    bits = bin(randbits(nbits))[2:].zfill(nbits)
    return Polynomial(lp=lp, coefs=decode2polycoefs(secpar, lp, distribution, dist_pars, bits, num_coefs, bti, btd),
                      const_time_flag=const_time_flag)


def random_polynomialvector(secpar: int, lp: LatticeParameters, distribution: str, dist_pars: Dict[str, int],
                            num_coefs: int, bti: int, btd: int, const_time_flag: bool = True) -> PolynomialVector:
    return PolynomialVector(lp=lp, entries=[
        random_polynomial(secpar, lp, distribution, dist_pars, num_coefs, bti, btd, const_time_flag)
        for _ in range(lp.length)], const_time_flag=const_time_flag)
