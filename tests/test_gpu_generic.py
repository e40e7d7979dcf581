"""Generic parameter sets: every power-of-two degree 32..1024 and NTT-friendly modulus runs on
ring_generic.cu / k_sampler_g instead of the d = 256 fast kernels.  Bit-exact checks through the C ABI against the
C oracle (schoolbook products, bit-by-bit decoder; oracle/lcb_oracle.c is generic in d) and against the restated
lattice_algebra; plus the reference's own container-test grid (tests/test_one_time_keys.py:12-33, the only
NTT-friendly pair it yields is (d, q) = (32, 193)) on the drop-in objects."""
import hashlib
from secrets import randbits

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (d, q, l): the reference's container grid, both directions from 256, moduli up to 16 bits
GRID = [(32, 193, 2), (64, 257, 3), (128, 769, 1), (512, 12289, 2), (1024, 12289, 2), (1024, 40961, 1), (512, 64513, 3)]


def _engine(secpar, q, d, l):
    from lattice_cryptography_b200 import Engine
    return Engine(secpar, q, d, l)


@pytest.mark.parametrize('d,q,l', GRID)
def test_context_and_root_of_unity(d, q, l):
    import lattice_algebra as ola                       # the restatement
    e = _engine(128, q, d, l)
    assert e.root_of_unity == ola.LatticeParameters(modulus=q, degree=d, length=l).rou
    e.close()


@pytest.mark.parametrize('d,q,l', GRID)
def test_transforms_and_product_vs_schoolbook(d, q, l):
    import c_oracle
    e = _engine(128, q, d, l)
    rng = np.random.default_rng(d + q)
    a = rng.integers(-(q // 2), q // 2 + 1, (9, d)).astype(np.int16)
    b = rng.integers(-(q // 2), q // 2 + 1, (9, d)).astype(np.int16)
    assert np.array_equal(e.ntt_inv(e.ntt_fwd(a)), a)
    got = e.poly_mul(a, b)
    for i in range(9):
        assert np.array_equal(got[i], c_oracle.poly_mul(q, a[i], b[i]))
    # any int16 input is reduced first (values beyond +-q/2 included)
    wild = rng.integers(-32768, 32768, (4, d)).astype(np.int16)
    cen = ((wild.astype(np.int64) + q // 2) % q - q // 2).astype(np.int16)
    assert np.array_equal(e.ntt_fwd(wild), e.ntt_fwd(cen))
    # addition / subtraction
    want = ((a.astype(np.int64) - b + q // 2) % q - q // 2).astype(np.int16)
    assert np.array_equal(e.vec_sub(a, b), want)
    e.close()


@pytest.mark.parametrize('d,q,l', [(32, 193, 2), (512, 12289, 2), (1024, 40961, 1)])
def test_reference_ntt_representation_l3(d, q, l):
    import lattice_algebra as ola
    e = _engine(128, q, d, l)
    lp = ola.LatticeParameters(modulus=q, degree=d, length=l)
    rng = np.random.default_rng(5)
    c = rng.integers(-(q // 2), q // 2 + 1, (3, d)).astype(np.int16)
    rep = e.ntt_reference_repr(c)
    assert rep.shape == (3, 2 * d)
    for i in range(3):
        p = ola.Polynomial(lp, {j: int(v) for j, v in enumerate(c[i]) if v}, False)
        assert rep[i].tolist() == [int(x) for x in p.ntt_representation]
    e.close()


@pytest.mark.parametrize('secpar', [128, 256])
@pytest.mark.parametrize('d,bd,wt,vec_len', [(32, 96, 32, 2), (32, 1, 1, 1), (32, 2, 2, 3), (64, 5, 17, 2), (512, 45, 512, 1),
                                             (512, 1, 20, 2), (1024, 65, 1024, 1), (1024, 6144, 300, 1), (128, 300, 128, 2)])
def test_sampler_vs_c_oracle(secpar, d, bd, wt, vec_len):
    import c_oracle
    q = {32: 193, 64: 257, 128: 769, 512: 12289, 1024: 12289}[d]
    e = _engine(secpar, q, d, 1)
    msgs = [b'', b'abc', bytes(range(200)) * 3, b'x' * 135, b'y' * 136]
    dense, pairs = e.hash2polyvec('SOME_SALT', msgs, bd, wt, vec_len, want_pairs=True)
    for i, m in enumerate(msgs):
        od, op = c_oracle.hash2polyvec(secpar, d, 'SOME_SALT', m, bd, wt, vec_len)
        assert np.array_equal(dense[i], od), (i, 'dense')
        assert np.array_equal(pairs[i], op), (i, 'draw order')
    e.close()


@pytest.mark.parametrize('d,q,l,sk_bd,ch_wt', [(32, 193, 3, 2, 4), (512, 12289, 3, 20, 30), (1024, 12289, 2, 4, 25),
                                               (1024, 40961, 2, 30, 60), (128, 769, 4, 3, 10)])
def test_lm_keygen_sign_verify_vs_c_oracle(d, q, l, sk_bd, ch_wt):
    """keygen / sign / verify parity at degrees 32 .. 1024 (lm_one_time_sigs.py:64-97,163-191)."""
    import c_oracle
    from lattice_cryptography_b200 import make_scheme
    secpar, n = 128, 21
    e = _engine(secpar, q, d, l)
    sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt)
    key_ch, _ = e.hash2polyvec('KEY_CH_SEED', ['generic ' + str(d)], q // 2, d, l)
    key_ch = np.ascontiguousarray(key_ch[0])
    e.set_key_ch(key_ch)
    op = c_oracle.params(secpar, q, l, sk_bd, ch_wt, d=d, sk_wt=d)
    seeds = [bin(1234567 * (j + 1))[2:].zfill(secpar) for j in range(n)]
    chm = [f'<key {j}>, ' + bin(99 + j)[2:].zfill(secpar - (j % 5)) for j in range(n)]
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
    sig = e.lm_sign(sch, sk_ntt, chm)
    vf_bd = min(q // 2, sk_bd * (1 + ch_wt))
    bad = sig.copy()
    bad[1, 0, 0] += 1 if bad[1, 0, 0] < vf_bd else -1          # breaks the equation
    bad[2, l - 1, d - 1] = vf_bd + 1                           # breaks the bound
    chm_bad = list(chm)
    chm_bad[3] = chm[3] + '!'
    verdict = e.lm_verify(sch, vk_ntt, chm_bad, bad, vf_bd, d)
    for j in range(n):
        skl, skr, vkl, vkr = c_oracle.lm_keygen(op, key_ch, seeds[j].encode())
        assert np.array_equal(sk_coef[j, 0], skl) and np.array_equal(sk_coef[j, 1], skr), j
        assert np.array_equal(vk_coef[j, 0], vkl) and np.array_equal(vk_coef[j, 1], vkr), j
        assert np.array_equal(sig[j], c_oracle.lm_sign(op, skl, skr, chm[j].encode())), j
        want = c_oracle.lm_verify(op, key_ch, vkl, vkr, chm_bad[j].encode(), bad[j], vf_bd, d)
        assert bool(verdict[j]) == want, j
    assert verdict.tolist() == [1, 0, 0, 0] + [1] * (n - 4)
    # the NTT-form keys are the transforms of the coefficient-form ones
    assert np.array_equal(e.ntt_inv(vk_ntt), vk_coef) and np.array_equal(e.ntt_inv(sk_ntt), sk_coef)
    e.close()


@pytest.mark.parametrize('d,q,l', [(32, 193, 2), (512, 12289, 2), (1024, 12289, 1)])
def test_bklm_and_adaptor_paths(d, q, l):
    """aggregation coefficients (first index = log2(d) bits), aggregate / aggregate_verify and the adaptor operations
    against the C oracle's primitives (bklm_one_time_agg_sigs.py:60-116, adaptor_sigs.py:80-101,191-266)."""
    import c_oracle
    from lattice_cryptography_b200 import make_scheme
    secpar, n, sk_bd, ch_wt = 128, 11, 2, 5
    e = _engine(secpar, q, d, l)
    sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt, wit_bd=1, wit_wt=7)
    key_ch, _ = e.hash2polyvec('KEY_CH_SEED', ['generic bklm ' + str(d)], q // 2, d, l)
    key_ch = np.ascontiguousarray(key_ch[0])
    e.set_key_ch(key_ch)
    op = c_oracle.params(secpar, q, l, sk_bd, ch_wt, d=d, sk_wt=d)
    seeds = [bin(777 * (j + 3))[2:].zfill(secpar) for j in range(n)]
    msgs = [bin(5 + j)[2:].zfill(32) for j in range(n)]
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * j:010x}>' for j in range(n)]
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    agmsg = ('[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']').encode()
    _, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds, want_sk_coef=False)
    sigs = e.lm_sign(sch, sk_ntt, chm)
    coefs = e.agg_coefs(sch, agmsg, 0, n)
    cen = lambda x: ((x.astype(np.int64) + q // 2) % q - q // 2).astype(np.int16)
    acc = np.zeros((l, d), dtype=np.int64)
    rhs = np.zeros(d, dtype=np.int64)
    for j in range(n):
        k, sgn = c_oracle.agg_coef(secpar, d, 'AG_SALT', j, agmsg)
        assert (int(coefs[j, 0, 0]), int(coefs[j, 0, 1])) == (k, sgn), j
        mono = np.zeros(d, dtype=np.int16)
        mono[k] = sgn
        for i in range(l):
            acc[i] += c_oracle.poly_mul(q, sigs[j, i], mono)
        ch, _ = c_oracle.hash2polyvec(secpar, d, 'CH_SALT', chm[j].encode(), 1, ch_wt, 1)
        t = cen(c_oracle.poly_mul(q, vk_coef[j, 0], ch[0]).astype(np.int64) + vk_coef[j, 1])
        rhs += c_oracle.poly_mul(q, t, mono)
    ag_sig = e.aggregate_finish(e.aggregate_partial(sch, sigs, coefs))
    assert np.array_equal(ag_sig, cen(acc))
    vpart = e.aggverify_partial(sch, vk_ntt, chm, coefs)
    assert np.array_equal(e.ntt_inv((vpart.astype(np.int64) % q).astype(np.uint16)[None])[0], cen(rhs))
    avf_bd = q // 2
    assert e.aggverify_finish(vpart, ag_sig, n, n, avf_bd, d) is True
    bad = ag_sig.copy()
    bad[0, 1] += 1 if bad[0, 1] < q // 2 else -1
    assert e.aggverify_finish(vpart, bad, n, n, avf_bd, d) is False
    assert e.aggverify_finish(vpart, ag_sig, n, n - 1, avf_bd, d) is False
    # adaptor: witness, statement, presign -> adapt -> verify -> extract -> witness_verify
    wit, st_ntt, st_coef = e.witgen(sch, seeds)
    vf_bd = min((q - 1) // 2, sk_bd * (1 + ch_wt))
    presig = sigs                                       # presign is sign over the adaptor's hash input
    full = e.vec_add(presig, wit)
    assert e.lm_verify(sch, vk_ntt, chm, presig, vf_bd, d).all()
    v = e.lm_verify(sch, vk_ntt, chm, full, vf_bd + 1, d, st_ntt=st_ntt)
    ext = e.vec_sub(full, presig)
    assert np.array_equal(ext, wit)
    wv = e.witness_verify(ext, st_ntt, 2 * vf_bd + 1, d)
    for j in range(n):
        ow, _ = c_oracle.hash2polyvec(secpar, d, 'WIT_SALT', seeds[j].encode(), 1, 7, l)
        assert np.array_equal(wit[j], ow), j
        assert np.array_equal(st_coef[j], c_oracle.dot(q, key_ch, ow)), j
        _, _, vkl, vkr = c_oracle.lm_keygen(op, key_ch, seeds[j].encode())
        assert bool(v[j]) == c_oracle.lm_verify(op, key_ch, vkl, vkr, chm[j].encode(), full[j], vf_bd + 1, d, st=st_coef[j])
    assert v.all() and wv.all()
    assert not e.lm_verify(sch, vk_ntt, chm, presig, vf_bd + 1, d, st_ntt=st_ntt).any()      # un-adapted
    e.close()


@pytest.mark.parametrize('d,q,l,ag_wt,ag_bd', [(256, 11777, 13, 3, 2), (256, 11777, 13, 1, 4), (256, 39937, 23, 2, 1),
                                               (512, 12289, 2, 4, 3), (32, 193, 2, 5, 1)])
def test_non_monomial_aggregation_coefficients(d, q, l, ag_wt, ag_bd):
    """ag_wt > 1 / ag_bd > 1 (bklm_one_time_agg_sigs.py:15-19 leaves both as editable tables; :96,114-115 use full
    polynomial products): coefficients against the C oracle's hash2polynomial with the salt 'AG_SALT' + str(i),
    aggregate and aggregate_verify against schoolbook products - on the shipped geometry too."""
    import c_oracle
    from lattice_cryptography_b200 import make_scheme
    secpar = 128 if q != 39937 else 256
    n, sk_bd, ch_wt = 9, 3, 6
    e = _engine(secpar, q, d, l)
    sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt, ag_bd=ag_bd, ag_wt=ag_wt)
    key_ch, _ = e.hash2polyvec('KEY_CH_SEED', ['non-monomial ' + str(d)], q // 2, d, l)
    key_ch = np.ascontiguousarray(key_ch[0])
    e.set_key_ch(key_ch)
    seeds = [bin(31337 * (j + 1))[2:].zfill(secpar) for j in range(n)]
    msgs = [bin(900 + j)[2:].zfill(32) for j in range(n)]
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * j:010x}>' for j in range(n)]
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    agmsg = ('[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']').encode()
    _, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds, want_sk_coef=False)
    sigs = e.lm_sign(sch, sk_ntt, chm)
    first = 7                                            # a shard that starts inside the sorted list
    coefs = e.agg_coefs(sch, agmsg, first, n)
    assert coefs.shape == (n, ag_wt, 2)
    cen = lambda x: ((x.astype(np.int64) + q // 2) % q - q // 2).astype(np.int16)
    acc = np.zeros((l, d), dtype=np.int64)
    rhs = np.zeros(d, dtype=np.int64)
    for j in range(n):
        dense, pairs = c_oracle.hash2polyvec(secpar, d, 'AG_SALT' + str(first + j), agmsg, ag_bd, ag_wt, 1)
        assert np.array_equal(coefs[j], pairs[0]), j
        for i in range(l):
            acc[i] += c_oracle.poly_mul(q, sigs[j, i], dense[0])
        ch, _ = c_oracle.hash2polyvec(secpar, d, 'CH_SALT', chm[j].encode(), 1, ch_wt, 1)
        t = cen(c_oracle.poly_mul(q, vk_coef[j, 0], ch[0]).astype(np.int64) + vk_coef[j, 1])
        rhs += c_oracle.poly_mul(q, t, dense[0])
    ag_sig = e.aggregate_finish(e.aggregate_partial(sch, sigs, coefs))
    assert np.array_equal(ag_sig, cen(acc))
    vpart = e.aggverify_partial(sch, vk_ntt, chm, coefs)
    assert np.array_equal(e.ntt_inv((vpart.astype(np.int64) % q).astype(np.uint16)[None])[0], cen(rhs))
    assert e.aggverify_finish(vpart, ag_sig, n, n, q // 2, d) is True
    bad = ag_sig.copy()
    bad[l - 1, d - 1] += 1 if bad[l - 1, d - 1] < q // 2 else -1
    assert e.aggverify_finish(vpart, bad, n, n, q // 2, d) is False
    e.close()


def test_wire_format_is_refused_off_the_shipped_geometry():
    from lattice_cryptography_b200 import LcbError
    e = _engine(128, 193, 32, 2)
    with pytest.raises((LcbError, ValueError)):
        e.pack(np.zeros((2, 32), dtype=np.int16), 8, 96)
    e.close()


def test_reference_container_grid_on_dropin_objects(monkeypatch):
    """The reference's tests/test_one_time_keys.py:36-239 on its own parameter grid - (d, q) = (32, 193), lengths 1..3,
    both secpars, witness bounds / weights 1..2 - run against the drop-in objects."""
    from lattice_cryptography_b200 import one_time_keys as otk
    from lattice_cryptography_b200.lattice_algebra import (LatticeParameters, Polynomial, PolynomialVector,
                                                           UNIFORM_INFINITY_WEIGHT, is_ntt_friendly_prime,
                                                           random_polynomialvector)
    pairs = [(2 ** k, q) for k in (5, 6, 7) for q in range(2 ** (k + 1) + 1, 2 ** 8, 2 ** (k + 1))
             if is_ntt_friendly_prime(modulus=q, degree=2 ** k)]
    assert pairs == [(32, 193)]
    for secpar in otk.ALLOWABLE_SECPARS:
        for (d, q) in pairs:
            for length in (1, 2, 3):
                lp = LatticeParameters(modulus=q, degree=d, length=length)
                seed = bin(randbits(secpar))[2:].zfill(secpar)
                ss = otk.SecretSeed(secpar=secpar, lp=lp, seed=seed)
                assert (ss.secpar, ss.lp, ss.seed) == (secpar, lp, seed) and ss == otk.SecretSeed(secpar=secpar, lp=lp, seed=seed)

                def rnd(bd, wt):
                    return random_polynomialvector(
                        secpar=secpar, lp=lp, distribution=UNIFORM_INFINITY_WEIGHT, dist_pars={'bd': bd, 'wt': wt},
                        num_coefs=wt, bti=otk.bits_per_index_set(secpar=secpar, degree=d, wt=wt),
                        btd=otk.bits_per_coefficient(secpar=secpar, bd=bd), const_time_flag=False)
                for bd in (1, 2):
                    for wt in (1, 2):
                        wit = otk.OneTimeSecretWitness(secpar=secpar, lp=lp, key=rnd(bd, wt))
                        assert isinstance(wit.key, PolynomialVector) and all(k.const_time_flag for k in wit.key.entries)
                        cnw = wit.key.get_coef_rep()
                        assert max(i[1] for i in cnw) <= bd and max(i[2] for i in cnw) <= wt
                        assert 1 <= min(i[1] for i in cnw) and min(i[2] for i in cnw) == wt
                        assert otk.OneTimeSecretWitness(secpar=secpar, lp=lp, key=wit.key) == wit
                        key_ch = rnd(q // 2, d)
                        stat = otk.OneTimePublicStatement(secpar=secpar, lp=lp, key=key_ch * wit.key)
                        assert isinstance(stat.key, Polynomial) and key_ch * wit.key == stat.key
                        assert stat == otk.OneTimePublicStatement(secpar=secpar, lp=lp, key=key_ch * wit.key)
                        left, right = rnd(q // 2, d), rnd(q // 2, d)
                        sk = otk.OneTimeSigningKey(secpar=secpar, lp=lp, left_key=left, right_key=right)
                        assert sk.left_key == left and sk.right_key == right
                        assert sk == otk.OneTimeSigningKey(secpar=secpar, lp=lp, left_key=left, right_key=right)
                        lvk, rvk = key_ch * sk.left_key, key_ch * sk.right_key
                        vk = otk.OneTimeVerificationKey(secpar=secpar, lp=lp, left_key=lvk, right_key=rvk)
                        assert vk[0] == lvk == vk.left_key and vk[1] == rvk == vk.right_key
                        assert vk == otk.OneTimeVerificationKey(secpar=secpar, lp=lp, left_key=lvk, right_key=rvk)
                        monkeypatch.setattr(otk, 'random_polynomialvector', lambda **kw: key_ch)
                        sp = otk.SchemeParameters(secpar=secpar, lp=lp, distribution=UNIFORM_INFINITY_WEIGHT, key_ch=None)
                        monkeypatch.undo()
                        assert (sp.secpar, sp.lp, sp.distribution) == (secpar, lp, UNIFORM_INFINITY_WEIGHT) and sp.key_ch == key_ch


def test_dropin_lm_scheme_at_other_parameters():
    """The scheme functions only read pp: a hand-made parameter set at d = 512 signs and verifies through the
    drop-in API (keygen -> sign -> verify, tampered message rejected)."""
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    from lattice_cryptography_b200.lattice_algebra import LatticeParameters
    from lattice_cryptography_b200.one_time_keys import SchemeParameters
    lp = LatticeParameters(modulus=12289, degree=512, length=4)
    sp = SchemeParameters(secpar=128, lp=lp, distribution=lm.DISTRIBUTION)
    pp = {'scheme_parameters': sp, 'sk_salt': 'SK_SALT', 'sk_bd': 10, 'sk_wt': 512, 'ch_salt': 'CH_SALT', 'ch_bd': 1,
          'ch_wt': 30}
    pp['vf_wt'] = max(1, min(lp.degree, pp['sk_wt'] * (1 + pp['ch_wt'])))
    pp['vf_bd'] = max(1, min(lp.modulus // 2, pp['sk_bd'] * (1 + min(pp['sk_wt'], pp['ch_wt']) * pp['ch_bd'])))
    keys = lm.keygen(pp=pp, num_keys_to_gen=3)
    for key in keys:
        assert sp.key_ch * key[1][0] == key[2][0] and sp.key_ch * key[1][1] == key[2][1]
        sig = lm.sign(pp=pp, otk=key, msg='QRL is awesome!')
        c = lm.make_signature_challenge(pp=pp, otvk=key[2], msg='QRL is awesome!')
        assert sig == key[1][0] ** c + key[1][1]
        assert sp.key_ch * sig == key[2][0] * c + key[2][1]
        assert lm.verify(pp=pp, otvk=key[2], msg='QRL is awesome!', sig=sig)
        assert not lm.verify(pp=pp, otvk=key[2], msg='QRL is awesome?', sig=sig)
