"""One-time adaptor signatures: the reference's entry points
(lattice_cryptography/adaptor_sigs.py:37-266) on the CUDA engine, plus batched variants.

Drop-in: make_setup_parameters, make_random_seed, make_one_wit, make_one_key, witgen, keygen,
         make_signature_challenge, presign, preverify, adapt, extract, witness_verify, sign, verify
Batched: witgen_batch, presign_batch, preverify_batch, adapt_batch, extract_batch,
         witness_verify_batch, verify_batch, challenge_messages  (keygen_batch is the LM one)
"""
from secrets import randbelow
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

from . import lm_one_time_sigs as _lm
from .lattice_algebra import LatticeParameters, Polynomial, PolynomialVector, hash2polynomial
from .lm_one_time_sigs import _ctx, keygen_batch  # noqa: F401
from .one_time_keys import (ALLOWABLE_SECPARS, OneTimePublicStatement, OneTimeSecretWitness, OneTimeSigningKey,
                            OneTimeVerificationKey, SchemeParameters, SecretSeed, UNIFORM_INFINITY_WEIGHT,
                            bits_to_decode, bits_to_indices)

SecurityParameter = int
PublicParameters = Dict[str, Any]
OneTimeKeyTuple = Tuple[SecretSeed, OneTimeSigningKey, OneTimeVerificationKey]
OneTimeWitStatPair = Tuple[SecretSeed, OneTimeSecretWitness, OneTimePublicStatement]
Message = str
Challenge = Polynomial
PreSignature = PolynomialVector
Signature = PolynomialVector

# shipped parameter sets (reference adaptor_sigs.py:19-34)
LPs: Dict[int, LatticeParameters] = dict(_lm.LPs)
BDs: Dict[int, Dict[str, int]] = {128: {'sk_bd': 45, 'ch_bd': 1, 'wit_bd': 1}, 256: {'sk_bd': 65, 'ch_bd': 1, 'wit_bd': 1}}
WTs: Dict[int, Dict[str, int]] = {128: {'sk_wt': 256, 'ch_wt': 20, 'wit_wt': 20},
                                  256: {'sk_wt': 256, 'ch_wt': 50, 'wit_wt': 20}}
SALTs: Dict[int, Dict[str, str]] = {i: {'sk_salt': 'SK_SALT', 'ch_salt': 'CH_SALT', 'wit_salt': 'WIT_SALT'}
                                    for i in ALLOWABLE_SECPARS}
DISTRIBUTION: str = UNIFORM_INFINITY_WEIGHT


def make_setup_parameters(secpar: SecurityParameter) -> PublicParameters:
    sp = SchemeParameters(secpar=secpar, lp=LPs[secpar], distribution=DISTRIBUTION)
    d, half = sp.lp.degree, (sp.lp.modulus - 1) // 2          # note (q-1)//2 here vs q//2 in the LM module
    pp: PublicParameters = {'scheme_parameters': sp}
    pp['sk_salt'], pp['sk_bd'], pp['sk_wt'] = SALTs[secpar]['sk_salt'], BDs[secpar]['sk_bd'], min(d, WTs[secpar]['sk_wt'])
    pp['ch_salt'], pp['ch_bd'], pp['ch_wt'] = SALTs[secpar]['ch_salt'], BDs[secpar]['ch_bd'], min(d, WTs[secpar]['ch_wt'])
    pp['wit_salt'], pp['wit_bd'], pp['wit_wt'] = SALTs[secpar]['wit_salt'], BDs[secpar]['wit_bd'], min(d, WTs[secpar]['wit_wt'])
    base_bd = pp['sk_bd'] * (1 + min(d, pp['sk_wt'], pp['ch_wt']) * pp['ch_bd'])
    pp['pvf_wt'] = max(1, min(d, pp['sk_wt'] * (1 + pp['ch_wt'])))
    pp['pvf_bd'] = max(1, min(half, base_bd))
    pp['vf_wt'] = max(1, min(d, pp['sk_wt'] * (1 + pp['ch_wt']) + pp['wit_wt']))
    pp['vf_bd'] = max(1, min(half, base_bd + pp['wit_bd']))
    pp['ext_wit_wt'] = max(1, min(d, pp['vf_wt'] + pp['pvf_wt']))
    pp['ext_wit_bd'] = max(1, min(half, pp['vf_bd'] + pp['pvf_bd']))
    return pp


def challenge_messages(sts: Sequence[Any], otvks: Sequence[Any], msgs: Sequence[Message]) -> List[str]:
    """Hash inputs of the adaptor challenge: str(st) + ', ' + str(otvk) + ', ' + msg (adaptor_sigs.py:176)."""
    return [str(s) + ', ' + str(k) + ', ' + m for s, k, m in zip(sts, otvks, msgs)]


# ------------------------------------------------------------------------------- batched API
def witgen_batch(pp: PublicParameters, seeds: Sequence[Any], device: bool = False):
    """-> dict(wit_coef int16[N,l,d], st_ntt uint16[N,d], st_coef int16[N,d])"""
    eng, sch = _ctx(pp)
    strs = [s.seed if isinstance(s, SecretSeed) else s for s in seeds]
    wit, st_ntt, st_coef = eng.witgen(sch, strs, device=device)
    return {'wit_coef': wit, 'st_ntt': st_ntt, 'st_coef': st_coef}


def presign_batch(pp: PublicParameters, sk_ntt, chmsgs, device: bool = False):
    eng, sch = _ctx(pp)
    return eng.lm_sign(sch, sk_ntt, chmsgs, device=device)


def preverify_batch(pp: PublicParameters, vk_ntt, chmsgs, presig, device: bool = False):
    eng, sch = _ctx(pp)
    return eng.lm_verify(sch, vk_ntt, chmsgs, presig, pp['pvf_bd'], pp['pvf_wt'], device=device)


def adapt_batch(pp: PublicParameters, presig, wit_coef, device: bool = False):
    return _ctx(pp)[0].vec_add(presig, wit_coef, device=device)


def extract_batch(pp: PublicParameters, presig, sig, device: bool = False):
    return _ctx(pp)[0].vec_sub(sig, presig, device=device)


def witness_verify_batch(pp: PublicParameters, wit_coef, st_ntt, device: bool = False):
    return _ctx(pp)[0].witness_verify(wit_coef, st_ntt, pp['ext_wit_bd'], pp['ext_wit_wt'], device=device)


def verify_batch(pp: PublicParameters, vk_ntt, chmsgs, st_ntt, sig, device: bool = False):
    eng, sch = _ctx(pp)
    return eng.lm_verify(sch, vk_ntt, chmsgs, sig, pp['vf_bd'], pp['vf_wt'], st_ntt=st_ntt, device=device)


# ------------------------------------------------------------------------------- drop-in API
def make_random_seed(secpar: SecurityParameter, pp: PublicParameters) -> SecretSeed:
    seed = bin(randbelow(2 ** secpar))[2:].zfill(secpar)
    return SecretSeed(secpar=secpar, lp=pp['scheme_parameters'].lp, seed=seed)


def _wrap_wits(pp: PublicParameters, seeds: List[SecretSeed], batch) -> List[OneTimeWitStatPair]:
    sp = pp['scheme_parameters']
    out = []
    for j, x in enumerate(seeds):
        wit = OneTimeSecretWitness(secpar=sp.secpar, lp=sp.lp,
                                   key=PolynomialVector(sp.lp, const_time_flag=True, _coef=batch['wit_coef'][j]))
        st = OneTimePublicStatement(secpar=sp.secpar, lp=sp.lp,
                                    key=Polynomial(sp.lp, _coef=batch['st_coef'][j], _ntt=batch['st_ntt'][j]))
        out.append((x, wit, st))
    return out


def _seed_list(pp: PublicParameters, count: int, seeds, what: str) -> List[SecretSeed]:
    if count < 1:
        raise ValueError(f'Can only generate a natural number worth of {what}.')
    elif seeds is not None and len(seeds) != count:
        raise ValueError(f'Must either roll {what} with no seeds, or with a seed for each key.')
    sp = pp['scheme_parameters']
    return list(seeds) if seeds is not None else [make_random_seed(secpar=sp.secpar, pp=pp) for _ in range(count)]


def make_one_wit(pp: PublicParameters, seed: SecretSeed = None) -> OneTimeWitStatPair:
    x = seed if seed else make_random_seed(secpar=pp['scheme_parameters'].secpar, pp=pp)
    return _wrap_wits(pp, [x], witgen_batch(pp, [x]))[0]


def make_one_key(pp: PublicParameters, seed: SecretSeed = None) -> OneTimeKeyTuple:
    return _lm.make_one_key(pp=pp, seed=seed)


def witgen(pp: PublicParameters, num_wits_to_gen: int = 1, seeds: List[SecretSeed] = None) -> List[OneTimeWitStatPair]:
    xs = _seed_list(pp, num_wits_to_gen, seeds, 'witnesses')
    return _wrap_wits(pp, xs, witgen_batch(pp, xs))


def keygen(pp: PublicParameters, num_keys_to_gen: int = 1, seeds: List[SecretSeed] = None) -> List[OneTimeKeyTuple]:
    xs = _seed_list(pp, num_keys_to_gen, seeds, 'keys')
    return _lm._wrap_keys(pp, xs, keygen_batch(pp, xs))


def make_signature_challenge(pp: PublicParameters, otvk: OneTimeVerificationKey, msg: Message,
                             st: OneTimePublicStatement) -> Challenge:
    sp = pp['scheme_parameters']
    return hash2polynomial(
        secpar=sp.secpar, lp=sp.lp, distribution=DISTRIBUTION, dist_pars={'bd': pp['ch_bd'], 'wt': pp['ch_wt']},
        salt=pp['ch_salt'], msg=str(st) + ', ' + str(otvk) + ', ' + msg, num_coefs=pp['ch_wt'],
        bti=bits_to_indices(secpar=sp.secpar, degree=sp.lp.degree, wt=pp['ch_wt']),
        btd=bits_to_decode(secpar=sp.secpar, bd=pp['ch_bd']), const_time_flag=True)


def _vk_arr(otvk: OneTimeVerificationKey) -> np.ndarray:
    return np.ascontiguousarray(np.stack([otvk[0].ntt, otvk[1].ntt])[None])


def presign(pp: PublicParameters, otk: OneTimeKeyTuple, msg: Message, st: OneTimePublicStatement) -> PreSignature:
    sk_ntt = np.ascontiguousarray(np.stack([otk[1][0].ntt, otk[1][1].ntt])[None])
    presig = presign_batch(pp, sk_ntt, challenge_messages([st], [otk[2]], [msg]))
    return PolynomialVector(pp['scheme_parameters'].lp, const_time_flag=True, _coef=presig[0])


def preverify(pp: PublicParameters, otvk: OneTimeVerificationKey, msg: Message, st: OneTimePublicStatement,
              presig: PreSignature) -> bool:
    if not _lm.well_formed(pp['scheme_parameters'].lp, presig):
        return False
    presig.const_time_flag = True
    verdict = preverify_batch(pp, _vk_arr(otvk), challenge_messages([st], [otvk], [msg]),
                              np.ascontiguousarray(presig.coef[None]))
    return bool(verdict[0])


def adapt(presig: PreSignature, wit: OneTimeSecretWitness) -> Signature:
    return presig + wit.key


def extract(pp: PublicParameters, presig: PreSignature, sig: Signature) -> OneTimeSecretWitness:
    sp = pp['scheme_parameters']
    return OneTimeSecretWitness(secpar=sp.secpar, lp=sp.lp, key=sig - presig)


def witness_verify(pp: PublicParameters, wit: OneTimeSecretWitness, st: OneTimePublicStatement) -> bool:
    if not _lm.well_formed(pp['scheme_parameters'].lp, wit.key):
        return False
    wit.const_time_flag = True
    verdict = witness_verify_batch(pp, np.ascontiguousarray(wit.key.coef[None]),
                                   np.ascontiguousarray(st.key.ntt[None]))
    return bool(verdict[0])


def sign(pp: PublicParameters, otk: OneTimeKeyTuple, msg: Message, wit_st_pair: OneTimeWitStatPair) -> Signature:
    presig = presign(pp=pp, otk=otk, msg=msg, st=wit_st_pair[2])
    return adapt(presig=presig, wit=wit_st_pair[1])


def verify(pp: PublicParameters, otvk: OneTimeVerificationKey, msg: Message, st: OneTimePublicStatement,
           sig: Signature) -> bool:
    if not _lm.well_formed(pp['scheme_parameters'].lp, sig):
        return False
    sig.const_time_flag = True
    verdict = verify_batch(pp, _vk_arr(otvk), challenge_messages([st], [otvk], [msg]),
                           np.ascontiguousarray(st.key.ntt[None]), np.ascontiguousarray(sig.coef[None]))
    return bool(verdict[0])
