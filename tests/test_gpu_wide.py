"""Moduli q >= 2^16 ("wide" contexts, SURVEY.md 8(f)4: "q up to 2^31"): every array of include/lcb200.h then holds
32-bit elements (int32 coefficients, uint32 NTT slots, int64 BKLM partial sums).  The C oracle is 16-bit only, so the
checker here is the restated lattice_algebra (oracle/lattice_algebra: 2d-point cyclic transform on Python integers)
and the restated scheme layer (oracle/schemes.py) with hand-made parameter dicts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (d, q, l): a Fermat prime, the Dilithium prime, the largest 31-bit NTT prime (15 * 2^27 + 1)
WIDE = [(256, 65537, 2), (64, 8380417, 3), (1024, 2013265921, 1), (32, 786433, 2)]


def _ctx(secpar, d, q, l):
    import lattice_algebra as ola
    from lattice_cryptography_b200 import Engine
    return Engine(secpar, q, d, l), ola.LatticeParameters(modulus=q, degree=d, length=l), ola


def _poly(ola, lp, dense):
    return ola.Polynomial(lp, {i: int(v) for i, v in enumerate(dense) if v}, False)


def _dense(p, d):
    out = np.zeros(d, dtype=np.int64)
    for i, v in p.get_coef_rep()[0].items():
        out[i] = v
    return out


@pytest.mark.parametrize('d,q,l', WIDE)
def test_wide_transforms_products_and_l3(d, q, l):
    e, lp, ola = _ctx(128, d, q, l)
    assert e.wide and e.ct is np.int32 and e.root_of_unity == lp.rou
    rng = np.random.default_rng(q % 1000 + d)
    a = rng.integers(-(q // 2), q // 2 + 1, (3, d)).astype(np.int32)
    b = rng.integers(-(q // 2), q // 2 + 1, (3, d)).astype(np.int32)
    fa = e.ntt_fwd(a)
    assert fa.dtype == np.uint32 and int(fa.max()) < q
    assert np.array_equal(e.ntt_inv(fa), a)
    got = e.poly_mul(a, b)
    for i in range(3):
        assert np.array_equal(got[i], _dense(_poly(ola, lp, a[i]) * _poly(ola, lp, b[i]), d)), i
    rep = e.ntt_reference_repr(a[:1])
    assert rep[0].tolist() == [int(x) for x in _poly(ola, lp, a[0]).ntt_representation]
    cen = lambda x: (x.astype(np.int64) + q // 2) % q - q // 2
    assert np.array_equal(e.vec_add(a, b), cen(a.astype(np.int64) + b))
    assert np.array_equal(e.vec_sub(a, b), cen(a.astype(np.int64) - b))
    # any int32 input is reduced first
    wild = rng.integers(-2 ** 31, 2 ** 31, (2, d)).astype(np.int32)
    assert np.array_equal(e.ntt_fwd(wild), e.ntt_fwd(cen(wild).astype(np.int32)))
    e.close()


@pytest.mark.parametrize('d,q,l,bd,wt', [(256, 65537, 2, 32768, 256), (64, 8380417, 3, 4190208, 64), (64, 8380417, 3, 70000, 9),
                                         (1024, 2013265921, 1, 1006632960, 40), (32, 786433, 2, 3, 5)])
def test_wide_sampler(d, q, l, bd, wt):
    """coefficient bounds beyond int16 (the public row key_ch has bd = q // 2): 64-bit field reduction in the decoder."""
    e, lp, ola = _ctx(128, d, q, l)
    msgs = ['', 'abc', 'z' * 300]
    dense, pairs = e.hash2polyvec('WIDE_SALT', msgs, bd, wt, l, want_pairs=True)
    assert dense.dtype == np.int32 and pairs.dtype == np.int32 and pairs.shape == (3, l, wt, 2)
    bti, btd = ola.bits_to_indices(128, d, wt), ola.bits_to_decode(128, bd)
    for i, m in enumerate(msgs):
        v = ola.hash2polynomialvector(secpar=128, lp=lp, distribution=ola.UNIFORM_INFINITY_WEIGHT, dist_pars={'bd': bd, 'wt': wt},
                                      num_coefs=wt, bti=bti, btd=btd, msg=m, salt='WIDE_SALT', const_time_flag=False)
        for j, p in enumerate(v.entries):
            assert np.array_equal(dense[i, j], _dense(p, d)), (i, j)
            assert {int(k): int(c) for k, c in pairs[i, j]} == p.get_coef_rep()[0]
    e.close()


@pytest.mark.parametrize('d,q,l,sk_bd,ch_wt', [(256, 65537, 2, 100, 20), (64, 8380417, 3, 40000, 8), (1024, 2013265921, 1, 5, 30)])
def test_wide_lm_bklm_adaptor_vs_restated_schemes(d, q, l, sk_bd, ch_wt):
    import schemes
    from lattice_cryptography_b200 import make_scheme
    secpar, n = 128, 3
    e, lp, ola = _ctx(secpar, d, q, l)
    sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt, wit_bd=1, wit_wt=6)
    key_ch_o = schemes.sample_vector(secpar, lp, 'KEY_CH_SEED', 'wide ' + str(q), q // 2, d)
    key_ch = np.array(schemes.dense_of_vec(key_ch_o), dtype=np.int32)
    got_kc, _ = e.hash2polyvec('KEY_CH_SEED', ['wide ' + str(q)], q // 2, d, l)
    assert np.array_equal(got_kc[0], key_ch)
    e.set_key_ch(key_ch)
    pp = dict(secpar=secpar, lp=lp, key_ch=key_ch_o, sk_salt='SK_SALT', ch_salt='CH_SALT', wit_salt='WIT_SALT', ag_salt='AG_SALT',
              sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt, wit_bd=1, wit_wt=6, ag_bd=1, ag_wt=1, ag_cap=n)
    pp['vf_wt'] = d
    pp['vf_bd'] = min(q // 2, sk_bd * (1 + ch_wt))
    pp['avf_wt'], pp['avf_bd'] = d, min(q // 2, n * pp['vf_bd'])
    seeds = [bin(424242 * (j + 1))[2:].zfill(secpar) for j in range(n)]
    msgs = [bin(77 + j)[2:].zfill(32) for j in range(n)]
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * j:010x}>' for j in range(n)]
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    agmsg = '[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']'
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
    sig = e.lm_sign(sch, sk_ntt, chm)
    okeys, osigs = [], []
    for j in range(n):
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, seeds[j])
        assert np.array_equal(sk_coef[j, 0], np.array(schemes.dense_of_vec(skl)))
        assert np.array_equal(vk_coef[j, 0], np.array(schemes.dense_of_poly(vkl))) and \
            np.array_equal(vk_coef[j, 1], np.array(schemes.dense_of_poly(vkr)))
        osig = schemes.lm_sign(pp, skl, skr, chm[j])
        assert np.array_equal(sig[j], np.array(schemes.dense_of_vec(osig))), j
        okeys.append((vkl, vkr))
        osigs.append(osig)
    bad = sig.copy()
    bad[1, 0, 3] += 1
    bad[2, l - 1, 0] = pp['vf_bd'] + 1
    verdict = e.lm_verify(sch, vk_ntt, chm, bad, pp['vf_bd'], d)
    assert verdict.tolist() == [1, 0, 0]
    for j in range(n):
        assert bool(verdict[j]) == schemes.lm_verify(pp, okeys[j][0], okeys[j][1], chm[j], schemes.vec_from_dense(lp, bad[j].tolist()))
    # BKLM (int64 partial sums)
    coefs = e.agg_coefs(sch, agmsg, 0, n)
    ocoefs = schemes.agg_coefs(pp, agmsg, n)
    for j in range(n):
        assert {int(coefs[j, 0, 0]): int(coefs[j, 0, 1])} == ocoefs[j].get_coef_rep()[0]
    part = e.aggregate_partial(sch, sig, coefs)
    assert part.dtype == np.int64
    ag_sig = e.aggregate_finish(part)
    oag = schemes.aggregate(pp, osigs, agmsg)
    assert np.array_equal(ag_sig, np.array(schemes.dense_of_vec(oag)))
    vpart = e.aggverify_partial(sch, vk_ntt, chm, coefs)
    assert e.aggverify_finish(vpart, ag_sig, n, n, pp['avf_bd'], d) is True
    assert schemes.aggregate_verify(pp, okeys, chm, agmsg, oag) is True
    tam = ag_sig.copy()
    tam[0, 0] += 1
    assert e.aggverify_finish(vpart, tam, n, n, pp['avf_bd'], d) is False
    # adaptor
    wit, st_ntt, st_coef = e.witgen(sch, seeds)
    for j in range(n):
        ow, ost = schemes.witgen_one(pp, seeds[j])
        assert np.array_equal(wit[j], np.array(schemes.dense_of_vec(ow))) and np.array_equal(st_coef[j], np.array(schemes.dense_of_poly(ost)))
    full = e.vec_add(sig, wit)
    assert e.lm_verify(sch, vk_ntt, chm, full, pp['vf_bd'] + 1, d, st_ntt=st_ntt).all()
    assert not e.lm_verify(sch, vk_ntt, chm, sig, pp['vf_bd'] + 1, d, st_ntt=st_ntt).any()
    assert np.array_equal(e.vec_sub(full, sig), wit)
    assert e.witness_verify(wit, st_ntt, 1, d).all()
    e.close()
