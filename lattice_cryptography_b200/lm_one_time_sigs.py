"""Lyubashevsky-Micciancio one-time signatures: the reference's entry points
(lattice_cryptography/lm_one_time_sigs.py:36-215) on the CUDA engine, plus batched variants.

Drop-in (same names, arguments, return types, error behaviour):
    make_setup_parameters, make_random_seed, make_one_key, keygen, keygen_core,
    make_signature_challenge, sign, verify, distribute_tasks
Batched (N independent keys / messages per call, arrays in engine format, see include/lcb200.h):
    keygen_batch, sign_batch, verify_batch, challenge_messages

`keygen(..., multiprocessing=...)` accepts the reference's keyword and ignores it: the whole batch
is one pass of the sampler + row-vector-product kernels instead of a process pool.
"""
from secrets import randbelow
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import make_scheme
from .lattice_algebra import (BoundEngine, LatticeParameters, Polynomial, PolynomialVector, engine_for, ensure_key_ch,
                              hash2polynomial, hash2polynomialvector)
from .one_time_keys import (ALLOWABLE_SECPARS, OneTimeSigningKey, OneTimeVerificationKey, SchemeParameters,
                            SecretSeed, UNIFORM_INFINITY_WEIGHT, bits_to_decode, bits_to_indices)

SecurityParameter = int
PublicParameters = Dict[str, Any]
OneTimeKeyTuple = Tuple[SecretSeed, OneTimeSigningKey, OneTimeVerificationKey]
Message = str
Challenge = Polynomial
Signature = PolynomialVector

# shipped parameter sets (reference lm_one_time_sigs.py:19-33)
LPs: Dict[int, LatticeParameters] = {
    128: LatticeParameters(modulus=11777, degree=2 ** 8, length=13),
    256: LatticeParameters(modulus=39937, degree=2 ** 8, length=23),
}
BDs: Dict[int, Dict[str, int]] = {128: {'sk_bd': 45, 'ch_bd': 1}, 256: {'sk_bd': 65, 'ch_bd': 1}}
WTs: Dict[int, Dict[str, int]] = {128: {'sk_wt': 256, 'ch_wt': 20}, 256: {'sk_wt': 256, 'ch_wt': 50}}
SALTs: Dict[int, Dict[str, str]] = {i: {'sk_salt': 'SK_SALT', 'ch_salt': 'CH_SALT'} for i in ALLOWABLE_SECPARS}
DISTRIBUTION: str = UNIFORM_INFINITY_WEIGHT


def make_setup_parameters(secpar: SecurityParameter) -> PublicParameters:
    sp = SchemeParameters(secpar=secpar, lp=LPs[secpar], distribution=DISTRIBUTION)
    pp: PublicParameters = {'scheme_parameters': sp}
    pp['sk_salt'], pp['sk_bd'], pp['sk_wt'] = SALTs[secpar]['sk_salt'], BDs[secpar]['sk_bd'], WTs[secpar]['sk_wt']
    pp['ch_salt'], pp['ch_bd'], pp['ch_wt'] = SALTs[secpar]['ch_salt'], BDs[secpar]['ch_bd'], WTs[secpar]['ch_wt']
    pp['vf_wt'] = max(1, min(sp.lp.degree, pp['sk_wt'] * (1 + pp['ch_wt'])))
    pp['vf_bd'] = max(1, min(sp.lp.modulus // 2, pp['sk_bd'] * (1 + min(pp['sk_wt'], pp['ch_wt']) * pp['ch_bd'])))
    return pp


# ------------------------------------------------------------------------------- engine plumbing
def _ctx(pp: PublicParameters):
    """(engine bound to pp's key_ch, lcb_scheme) for a pp dict; the row is (re)uploaded inside each call whenever the
    cached context holds a different one."""
    sp = pp['scheme_parameters']
    eng = BoundEngine(engine_for(sp.lp, sp.secpar), sp.key_ch.coef)
    sch = make_scheme(sk_bd=pp.get('sk_bd', 1), sk_wt=pp.get('sk_wt', 1), ch_bd=pp.get('ch_bd', 1),
                      ch_wt=pp.get('ch_wt', 1), ag_bd=pp.get('ag_bd', 1), ag_wt=pp.get('ag_wt', 1),
                      wit_bd=pp.get('wit_bd', 1), wit_wt=pp.get('wit_wt', 1), sk_salt=pp.get('sk_salt', 'SK_SALT'),
                      ch_salt=pp.get('ch_salt', 'CH_SALT'), ag_salt=pp.get('ag_salt', 'AG_SALT'),
                      wit_salt=pp.get('wit_salt', 'WIT_SALT'))
    return eng, sch


def challenge_messages(otvks: Sequence[Any], msgs: Sequence[Message]) -> List[str]:
    """The hash input of make_signature_challenge for each (otvk, msg): str(otvk) + ', ' + msg
    (reference lm_one_time_sigs.py:148).  str() of a key object is CPython's address-based default,
    so a signature verifies only against the same live key object - as in the reference."""
    return [str(k) + ', ' + m for k, m in zip(otvks, msgs)]


# ------------------------------------------------------------------------------- batched API
def keygen_batch(pp: PublicParameters, seeds: Sequence[Any], want_coef: bool = True, device: bool = False):
    """N keys from N seeds (SecretSeed objects or bitstrings) in one engine call.
    -> dict(sk_ntt uint16[N,2,l,d], vk_ntt uint16[N,2,d][, sk_coef int16[N,2,l,d], vk_coef int16[N,2,d]])"""
    eng, sch = _ctx(pp)
    if isinstance(seeds, tuple) and len(seeds) == 2 and not isinstance(seeds[0], (str, SecretSeed)):
        strs = seeds                # already (byte blob, int64 offsets), e.g. from random_seed_batch
    else:
        strs = [s.seed if isinstance(s, SecretSeed) else s for s in seeds]
    sk_coef, sk_ntt, vk_ntt, vk_coef = eng.lm_keygen(sch, strs, want_sk_coef=want_coef, want_vk_coef=want_coef,
                                                     device=device)
    return {'sk_ntt': sk_ntt, 'vk_ntt': vk_ntt, 'sk_coef': sk_coef, 'vk_coef': vk_coef}


def random_seed_batch(pp: PublicParameters, n: int, device: bool = False, secret: Optional[bytes] = None):
    """The unseeded keygen path (make_random_seed, reference lm_one_time_sigs.py:58-61) for a batch: ONE fresh 32-byte
    secret from the OS entropy source (`secrets`) is expanded on the GPU (SHAKE256 in counter mode, lcb_expand_seeds)
    into n secpar-bit seeds, as the ASCII bitstrings the reference hashes, in the (blob, offsets) form keygen_batch
    takes.  Row i of blob.reshape(n, secpar) is seed i: it reproduces key i through the seeded path.  `secret`
    (32 bytes) makes the batch reproducible; by default nobody ever sees it."""
    from secrets import token_bytes
    eng, _ = _ctx(pp)
    return eng.expand_seeds(secret if secret is not None else token_bytes(32), n, device=device)


def sign_batch(pp: PublicParameters, sk_ntt, chmsgs, device: bool = False):
    """N signatures: sig[i] = sk_left[i] ** c_i + sk_right[i], c_i = H(ch_salt || chmsgs[i]).
    sk_ntt uint16[N,2,l,d] (from keygen_batch) -> int16[N,l,d] centred coefficients."""
    eng, sch = _ctx(pp)
    return eng.lm_sign(sch, sk_ntt, chmsgs, device=device)


def verify_batch(pp: PublicParameters, vk_ntt, chmsgs, sig, device: bool = False):
    """N verdicts (uint8): norm / weight bounds and key_ch * sig == vk_left * c + vk_right."""
    eng, sch = _ctx(pp)
    return eng.lm_verify(sch, vk_ntt, chmsgs, sig, pp['vf_bd'], pp['vf_wt'], device=device)


# ------------------------------------------------------------------------------- drop-in API
def make_random_seed(secpar: SecurityParameter, pp: PublicParameters) -> SecretSeed:
    seed = bin(randbelow(2 ** secpar))[2:].zfill(secpar)
    return SecretSeed(secpar=secpar, lp=pp['scheme_parameters'].lp, seed=seed)


def _wrap_keys(pp: PublicParameters, seeds: List[SecretSeed], batch) -> List[OneTimeKeyTuple]:
    sp = pp['scheme_parameters']
    out = []
    for j, x in enumerate(seeds):
        halves = [PolynomialVector(sp.lp, const_time_flag=True, _coef=batch['sk_coef'][j, h], _ntt=batch['sk_ntt'][j, h])
                  for h in (0, 1)]
        vks = [Polynomial(sp.lp, const_time_flag=False, _coef=batch['vk_coef'][j, h], _ntt=batch['vk_ntt'][j, h])
               for h in (0, 1)]
        otsk = OneTimeSigningKey(secpar=sp.secpar, lp=sp.lp, left_key=halves[0], right_key=halves[1])
        otvk = OneTimeVerificationKey(secpar=sp.secpar, lp=sp.lp, left_key=vks[0], right_key=vks[1])
        out.append((x, otsk, otvk))
    return out


def make_one_key(pp: PublicParameters, seed: SecretSeed = None) -> OneTimeKeyTuple:
    sp = pp['scheme_parameters']
    x = seed if seed else make_random_seed(secpar=sp.secpar, pp=pp)
    sp.key_ch.const_time_flag = True
    return _wrap_keys(pp, [x], keygen_batch(pp, [x]))[0]


def keygen_core(pp: PublicParameters, num_keys_to_gen: int = 1, seeds: List[SecretSeed] = None) -> List[OneTimeKeyTuple]:
    if num_keys_to_gen < 1:
        raise ValueError('Can only generate a natural number worth of keys.')
    elif seeds is not None and len(seeds) != num_keys_to_gen:
        raise ValueError('Must either roll keys with no seeds, or with a seed for each key.')
    sp = pp['scheme_parameters']
    xs = list(seeds) if seeds is not None else [make_random_seed(secpar=sp.secpar, pp=pp) for _ in range(num_keys_to_gen)]
    sp.key_ch.const_time_flag = True
    return _wrap_keys(pp, xs, keygen_batch(pp, xs))


def keygen(pp: PublicParameters, num_keys_to_gen: int = 1, seeds: List[SecretSeed] = None,
           multiprocessing: bool = None) -> List[OneTimeKeyTuple]:
    """Same contract as the reference's Pool wrapper (lm_one_time_sigs.py:100-123), order preserved; the batch is
    one pass of the sampler + row-vector-product kernels instead of a process pool.  The observable corner
    cases of the wrapper are kept: without "multiprocessing" (fewer than 16 keys by default) the arguments go to
    keygen_core unchanged, so an empty seed list raises; with it, a falsy seed list means "no seeds" and a seed
    list of another length yields min(len(seeds), num_keys_to_gen) keys."""
    if multiprocessing is None:
        multiprocessing = num_keys_to_gen >= 16
    if (not multiprocessing) or num_keys_to_gen == 1:
        return keygen_core(pp=pp, num_keys_to_gen=num_keys_to_gen, seeds=seeds)
    if not seeds:
        return keygen_core(pp=pp, num_keys_to_gen=num_keys_to_gen, seeds=None)
    return keygen_core(pp=pp, num_keys_to_gen=len(seeds), seeds=seeds)[:num_keys_to_gen]


def make_signature_challenge(pp: PublicParameters, otvk: OneTimeVerificationKey, msg: Message) -> Challenge:
    sp = pp['scheme_parameters']
    return hash2polynomial(
        secpar=sp.secpar, lp=sp.lp, distribution=DISTRIBUTION, dist_pars={'bd': pp['ch_bd'], 'wt': pp['ch_wt']},
        salt=pp['ch_salt'], msg=str(otvk) + ', ' + msg, num_coefs=pp['ch_wt'],
        bti=bits_to_indices(secpar=sp.secpar, degree=sp.lp.degree, wt=pp['ch_wt']),
        btd=bits_to_decode(secpar=sp.secpar, bd=pp['ch_bd']), const_time_flag=True)


def sign(pp: PublicParameters, otk: OneTimeKeyTuple, msg: Message) -> Signature:
    sp = pp['scheme_parameters']
    sk_ntt = np.ascontiguousarray(np.stack([otk[1][0].ntt, otk[1][1].ntt])[None])
    sig = sign_batch(pp, sk_ntt, challenge_messages([otk[2]], [msg]))
    return PolynomialVector(sp.lp, const_time_flag=False, _coef=sig[0])


def well_formed(lp: LatticeParameters, vec) -> bool:
    """True iff `vec` is a PolynomialVector over `lp` with exactly lp.length entries of lp.degree coefficients.
    The reference's verify functions pair entries with zip() and so tolerate a short vector; the engine reads
    l*d coefficients from the buffer it is given, so a malformed signature is rejected here (verdict False)."""
    return isinstance(vec, PolynomialVector) and vec.lp == lp and tuple(vec.coef.shape) == (lp.length, lp.degree)


def verify(pp: PublicParameters, otvk: OneTimeVerificationKey, msg: Message, sig: Signature) -> bool:
    sp = pp['scheme_parameters']
    if not well_formed(sp.lp, sig):
        return False
    sig.const_time_flag = False
    sp.key_ch.const_time_flag = False
    otvk.left_key.const_time_flag = False
    otvk.right_key.const_time_flag = False
    vk_ntt = np.ascontiguousarray(np.stack([otvk[0].ntt, otvk[1].ntt])[None])
    verdict = verify_batch(pp, vk_ntt, challenge_messages([otvk], [msg]), np.ascontiguousarray(sig.coef[None]))
    return bool(verdict[0])


def distribute_tasks(tasks: List[Any], num_workers: int = None) -> List[List[Any]]:
    """Contiguous split of `tasks` into num_workers shards, the first len % num_workers one longer
    (reference lm_one_time_sigs.py:194-215); used here to shard batches over GPUs / ranks."""
    if not num_workers:
        from multiprocessing import cpu_count
        num_workers = cpu_count()
    base, extra = divmod(len(tasks), num_workers)
    out, start = [], 0
    for w in range(num_workers):
        size = base + (1 if w < extra else 0)
        out.append(tasks[start:start + size])
        start += size
    return out
