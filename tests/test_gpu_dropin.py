"""GPU tests of the drop-in Python layer: the reference's own integration tests and the algebraic
identities its unit tests assert, run against lattice_cryptography_b200's modules.
  * tests/test_bklm_one_time_agg_sigs.py:406-415 (test_all), tests/test_adaptor_sigs.py:196-217
    (test_general), benchmarks/demo_signing.py
  * tests/test_lm_one_time_sigs.py:164-176,284-292 identities; no-forgery checks the reference lacks.
Objects and verdicts are also compared with the CPU oracle on the same key_ch / seeds / hash inputs."""
from secrets import randbits

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('secpar', [128, 256])
def test_lm_dropin_identities(secpar):
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    from lattice_cryptography_b200.one_time_keys import SecretSeed
    pp = lm.make_setup_parameters(secpar)
    sp = pp['scheme_parameters']
    assert pp['vf_bd'] == {128: 945, 256: 3315}[secpar] and pp['vf_wt'] == 256
    seeds = [SecretSeed(secpar=secpar, lp=sp.lp, seed=bin(j)[2:].zfill(secpar)) for j in (0, 1, 12345)]
    keys = lm.keygen(pp=pp, num_keys_to_gen=3, seeds=seeds, multiprocessing=True)
    assert [k[0] for k in keys] == seeds
    for seed, sk, vk in keys:
        assert sk.left_key.const_time_flag and sk.right_key.const_time_flag
        assert not vk.left_key.const_time_flag and not vk.right_key.const_time_flag
        assert sp.key_ch * sk[0] == vk[0] == vk.left_key
        assert sp.key_ch * sk[1] == vk[1] == vk.right_key
        for entry in sk[0].get_coef_rep() + sk[1].get_coef_rep():
            assert 1 <= entry[1] <= pp['sk_bd'] and 1 <= entry[2] <= pp['sk_wt']
        msg = 'QRL is awesome!'
        c = lm.make_signature_challenge(pp=pp, otvk=vk, msg=msg)
        coefs, norm, weight = c.get_coef_rep()
        assert norm == 1 and weight == pp['ch_wt'] and set(coefs.values()) <= {-1, 1}
        sig = lm.sign(pp=pp, otk=(seed, sk, vk), msg=msg)
        assert sig == sk[0] ** c + sk[1]
        assert sp.key_ch * sig == vk[0] * c + vk[1]
        cnw = sig.get_coef_rep()
        assert 1 <= max(i[1] for i in cnw) <= pp['vf_bd'] and 1 <= max(i[2] for i in cnw) <= pp['vf_wt']
        assert lm.verify(pp=pp, otvk=vk, msg=msg, sig=sig) is True
        assert lm.verify(pp=pp, otvk=vk, msg=msg + ' ', sig=sig) is False
        assert lm.verify(pp=pp, otvk=keys[0][2] if vk is not keys[0][2] else keys[1][2], msg=msg, sig=sig) is False
    # unseeded path, single key and the error contract of keygen_core (lm_one_time_sigs.py:128-131)
    one = lm.keygen(pp=pp, num_keys_to_gen=1)[0]
    assert len(one[0].seed) == secpar and lm.verify(pp, one[2], '0101', lm.sign(pp, one, '0101'))
    with pytest.raises(ValueError, match='natural number'):
        lm.keygen_core(pp=pp, num_keys_to_gen=0)
    with pytest.raises(ValueError, match='seed for each key'):
        lm.keygen_core(pp=pp, num_keys_to_gen=2, seeds=seeds)


def test_content_str_mode_survives_pickling():
    """Opt-in content-based str(otvk) (one_time_keys.set_content_str): a signature then verifies against a
    pickled copy of the key; with the reference's address-based default it does not."""
    import pickle
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    from lattice_cryptography_b200 import one_time_keys as otk
    pp = lm.make_setup_parameters(128)
    key = lm.keygen(pp=pp, num_keys_to_gen=1)[0]
    sig = lm.sign(pp=pp, otk=key, msg='0110')
    assert lm.verify(pp=pp, otvk=pickle.loads(pickle.dumps(key[2])), msg='0110', sig=sig) is False
    otk.set_content_str(True)
    try:
        sig = lm.sign(pp=pp, otk=key, msg='0110')
        clone = pickle.loads(pickle.dumps(key[2]))
        assert lm.verify(pp=pp, otvk=clone, msg='0110', sig=sig) is True
        assert lm.verify(pp=pp, otvk=clone, msg='0111', sig=sig) is False
        # BKLM builds its aggregation message from repr() of the keys inside a list
        bpp = bk.make_setup_parameters(128)
        keys = lm.keygen(pp=bpp, num_keys_to_gen=2)
        msgs = ['0101', '1100']
        sigs = [lm.sign(pp=bpp, otk=k, msg=m) for k, m in zip(keys, msgs)]
        ag = bk.aggregate(pp=bpp, otvks=[k[2] for k in keys], msgs=msgs, sigs=sigs)
        clones = [pickle.loads(pickle.dumps(k[2])) for k in keys]
        assert bk.aggregate_verify(pp=bpp, otvks=clones, msgs=msgs, ag_sig=ag) is True
    finally:
        otk.set_content_str(False)


def test_lm_dropin_matches_oracle():
    """Same key_ch, seed and full hash input -> same keys, challenge and signature as the CPU oracle."""
    import schemes
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    from lattice_cryptography_b200.lattice_algebra import PolynomialVector
    from lattice_cryptography_b200.one_time_keys import SchemeParameters, SecretSeed
    secpar = 128
    okc = schemes.key_ch_from_seed(secpar, 'drop-in parity key_ch')
    opp = schemes.make_lm_parameters(secpar, okc)
    pp = lm.make_setup_parameters(secpar)
    lp = pp['scheme_parameters'].lp
    pp['scheme_parameters'] = SchemeParameters(secpar=secpar, lp=lp, distribution=lm.DISTRIBUTION, key_ch=PolynomialVector(
        lp, _coef=np.array(schemes.dense_of_vec(okc), dtype=np.int16)))
    seed = bin(987654321)[2:].zfill(secpar)
    key = lm.keygen(pp, 1, [SecretSeed(seed=seed, secpar=secpar, lp=lp)])[0]
    skl, skr, vkl, vkr = schemes.lm_keygen_one(opp, seed)
    assert key[1][0].coef.tolist() == schemes.dense_of_vec(skl) and key[1][1].coef.tolist() == schemes.dense_of_vec(skr)
    assert key[2][0].coef.tolist() == schemes.dense_of_poly(vkl) and key[2][1].coef.tolist() == schemes.dense_of_poly(vkr)
    msg = 'Blessed are the cheesemakers.'
    chmsg = str(key[2]) + ', ' + msg
    sig = lm.sign(pp, key, msg)
    assert sig.coef.tolist() == schemes.dense_of_vec(schemes.lm_sign(opp, skl, skr, chmsg))
    assert lm.make_signature_challenge(pp, key[2], msg).coef.tolist() == schemes.dense_of_poly(schemes.challenge(opp, chmsg))


@pytest.mark.parametrize('secpar', [128, 256])
def test_bklm_all(secpar):
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import keygen, sign, verify
    pp = bk.make_setup_parameters(secpar)
    assert pp['ag_cap'] == 2 and pp['avf_bd'] == {128: 1890, 256: 6630}[secpar] and pp['avf_wt'] == 256
    for _ in range(4):
        keys = keygen(pp=pp, num_keys_to_gen=pp['ag_cap'])
        msgs = [bin(randbits(32))[2:].zfill(32) for _ in keys]
        sigs = [sign(pp=pp, otk=k, msg=m) for k, m in zip(keys, msgs)]
        assert all(verify(pp=pp, otvk=k[2], msg=m, sig=s) for k, m, s in zip(keys, msgs, sigs))
        vks = [k[2] for k in keys]
        ag_sig = bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs)
        # the engine's aggregate equals the object-level formula of the reference (bklm_one_time_agg_sigs.py:92-96)
        srt_keys, srt_msgs, srt_sigs = bk.prepare_aggregate(otvks=vks, msgs=msgs, sigs=sigs)
        coefs = bk.make_agg_coefs(pp=pp, otvks=vks, msgs=msgs)
        assert all(c.get_coef_rep()[1:] == (1, 1) for c in coefs)
        assert ag_sig == sum([s ** c for s, c in zip(srt_sigs, coefs)])
        assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs, ag_sig=ag_sig)
        assert bk.aggregate_verify(pp=pp, otvks=vks[::-1], msgs=msgs[::-1], ag_sig=ag_sig)      # order-independent
        assert not bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs[::-1], ag_sig=ag_sig)
        assert not bk.aggregate_verify(pp=pp, otvks=vks[:1], msgs=msgs, ag_sig=ag_sig)          # length mismatch
        assert not bk.aggregate_verify(pp=pp, otvks=vks + vks[:1], msgs=msgs + msgs[:1], ag_sig=ag_sig)   # over cap
    with pytest.raises(ValueError, match='bitstrings'):
        bk.prepare_make_agg_coefs(vks, ['01', 'xy'])
    with pytest.raises(ValueError, match='equal length'):
        bk.prepare_make_agg_coefs(vks, ['01'])


def test_bklm_larger_aggregate_sharded():
    """ag_cap raised to 24 and the sorted list split into 3 shards, as 3 ranks would."""
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import challenge_messages, keygen, sign
    pp = bk.set_aggregation_capacity(bk.make_setup_parameters(128), 24)
    assert pp['avf_bd'] == 5888 and pp['avf_wt'] == 256
    keys = keygen(pp=pp, num_keys_to_gen=24)
    msgs = [bin(randbits(32))[2:].zfill(32) for _ in keys]
    sigs = [sign(pp=pp, otk=k, msg=m) for k, m in zip(keys, msgs)]
    vks = [k[2] for k in keys]
    whole = bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs)
    srt_keys, srt_msgs, srt_sigs = bk.prepare_aggregate(vks, msgs, sigs)
    agmsg = str(list(zip(srt_keys, srt_msgs)))
    sig_arr = np.stack([s.coef for s in srt_sigs])
    vk_arr = np.stack([np.stack([k[0].ntt, k[1].ntt]) for k in srt_keys])
    chm = challenge_messages(srt_keys, srt_msgs)
    part = np.zeros((13, 256), dtype=np.int64)
    vpart = np.zeros(256, dtype=np.int64)
    for a, b in ((0, 7), (7, 16), (16, 24)):
        part += bk.aggregate_shard(pp, np.ascontiguousarray(sig_arr[a:b]), agmsg, a)
        vpart += bk.aggregate_verify_shard(pp, np.ascontiguousarray(vk_arr[a:b]), chm[a:b], agmsg, a)
    ag = bk.aggregate_finish(pp, part.astype(np.int32))
    assert np.array_equal(ag, whole.coef)
    assert bk.aggregate_verify_finish(pp, vpart.astype(np.int32), ag, 24)
    assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs, ag_sig=whole)


def test_adaptor_general():
    from lattice_cryptography_b200 import adaptor_sigs as ad
    for secpar in (128, 256):
        pp = ad.make_setup_parameters(secpar=secpar)
        assert (pp['pvf_bd'], pp['vf_bd'], pp['ext_wit_bd']) == {128: (945, 946, 1891), 256: (3315, 3316, 6631)}[secpar]
        key = ad.keygen(pp=pp, num_keys_to_gen=1)[0]
        otvk = key[2]
        seed, wit, st = ad.witgen(pp=pp, num_wits_to_gen=1)[0]
        assert pp['scheme_parameters'].key_ch * wit.key == st.key
        message = 'Blessed are the cheesemakers.'
        presig = ad.presign(pp=pp, otk=key, msg=message, st=st)
        assert ad.preverify(pp=pp, otvk=otvk, msg=message, st=st, presig=presig)
        sig = ad.adapt(presig=presig, wit=wit)
        assert ad.verify(pp=pp, otvk=otvk, msg=message, st=st, sig=sig)
        assert not ad.verify(pp=pp, otvk=otvk, msg=message, st=st, sig=presig)        # un-adapted
        ext = ad.extract(pp=pp, sig=sig, presig=presig)
        assert ext.key == wit.key
        assert ad.witness_verify(pp=pp, wit=ext, st=st)
        signed = ad.sign(pp=pp, otk=key, msg=message, wit_st_pair=(seed, wit, st))
        assert signed == sig and ad.verify(pp=pp, otvk=otvk, msg=message, st=st, sig=signed)
        other = ad.witgen(pp=pp, num_wits_to_gen=1)[0]
        assert not ad.witness_verify(pp=pp, wit=other[1], st=st)
        with pytest.raises(ValueError, match='witnesses'):
            ad.witgen(pp=pp, num_wits_to_gen=0)


def test_container_validation_messages():
    """Error strings of the key containers (one_time_keys.py:26-38 etc.) - host logic, but the module
    imports the engine, so it lives with the GPU tests."""
    from lattice_cryptography_b200 import one_time_keys as otk
    from lattice_cryptography_b200.lm_one_time_sigs import LPs
    lp = LPs[128]
    with pytest.raises(ValueError, match=r'must be an integer in \[128, 256\] but had 127'):
        otk.SecretSeed(seed='0' * 128, secpar=127, lp=lp)
    with pytest.raises(ValueError, match='Input must be a binary string'):
        otk.SecretSeed(seed='012', secpar=128, lp=lp)
    with pytest.raises(ValueError, match='enough bits'):
        otk.SecretSeed(seed='01', secpar=128, lp=lp)
    assert otk.SecretSeed(seed='0' * 128, secpar=128, lp=lp) == otk.SecretSeed(seed='0' * 128, secpar=128, lp=lp)
    assert otk.bits_to_indices(128, 256, 20) == 2592 and otk.bits_to_decode(128, 1) == 129
    with pytest.raises(ValueError, match='non-positive'):
        otk.bits_per_coefficient(128, 0)


@pytest.mark.parametrize('secpar', [128, 256])
def test_bklm_all_at_the_reference_sample_size(secpar):
    """tests/test_bklm_one_time_agg_sigs.py:406-415 (test_all) as the reference runs it: SAMPLE_SIZE = 64 rounds per
    security parameter of keygen(ag_cap) -> sign -> verify -> aggregate -> aggregate_verify, fresh random keys and
    32-bit messages every round."""
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import keygen, sign, verify
    pp = bk.make_setup_parameters(secpar)
    for _ in range(64):
        keys = keygen(pp=pp, num_keys_to_gen=pp['ag_cap'])
        msgs = [bin(randbits(32))[2:].zfill(32) for _ in range(pp['ag_cap'])]
        sigs = [sign(pp=pp, otk=k, msg=m) for k, m in zip(keys, msgs)]
        for k, m, s in zip(keys, msgs, sigs):
            assert verify(pp=pp, otvk=k[2], msg=m, sig=s)
        otvks = [k[2] for k in keys]
        ag_sig = bk.aggregate(pp=pp, otvks=otvks, msgs=msgs, sigs=sigs)
        assert bk.aggregate_verify(pp=pp, otvks=otvks, msgs=msgs, ag_sig=ag_sig)


@pytest.mark.parametrize('secpar', [128, 256])
def test_lower_seam_alone_carries_the_schemes(secpar):
    """INTEGRATION.md option (b): the reference's scheme modules only use the ten names they import from
    lattice_algebra.  Here every scheme operation is computed a second time from those names
    and operators ALONE on the GPU-backed objects - the way the reference's own function bodies do it
    (lm_one_time_sigs.py:64-97,141-191; bklm_one_time_agg_sigs.py:60-116; adaptor_sigs.py:80-101,191-266) - and must
    agree with the batched drop-in functions object for object and verdict for verdict."""
    from lattice_cryptography_b200 import adaptor_sigs as ad
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    from lattice_cryptography_b200.lattice_algebra import (UNIFORM_INFINITY_WEIGHT as DIST, bits_to_decode, bits_to_indices,
                                                           hash2polynomial, hash2polynomialvector)
    pp = ad.make_setup_parameters(secpar)
    bpp = bk.make_setup_parameters(secpar)
    bpp['scheme_parameters'] = pp['scheme_parameters']
    sp = pp['scheme_parameters']
    lp, key_ch, d = sp.lp, sp.key_ch, sp.lp.degree

    def h2pv(salt, msg, bd, wt):
        return hash2polynomialvector(secpar=secpar, lp=lp, distribution=DIST, dist_pars={'bd': bd, 'wt': wt}, num_coefs=wt,
                                     bti=bits_to_indices(secpar, d, wt), btd=bits_to_decode(secpar, bd), msg=msg, salt=salt,
                                     const_time_flag=False)

    def h2p(salt, msg, bd, wt):
        return hash2polynomial(secpar=secpar, lp=lp, distribution=DIST, dist_pars={'bd': bd, 'wt': wt}, salt=salt, msg=msg,
                               num_coefs=wt, bti=bits_to_indices(secpar, d, wt), btd=bits_to_decode(secpar, bd),
                               const_time_flag=False)

    def within(vec, bd, wt, lower=False):
        cnw = vec.get_coef_rep()
        n, w = max(i[1] for i in cnw), max(i[2] for i in cnw)
        return n <= bd and w <= wt and (not lower or (n >= 1 and w >= 1))

    keys = lm.keygen(pp=pp, num_keys_to_gen=2)
    msgs = [bin(randbits(32))[2:].zfill(32) for _ in keys]
    for (seed, sk, vk), msg in zip(keys, msgs):
        # make_one_key
        left, right = h2pv(pp['sk_salt'] + 'LEFT', seed.seed, pp['sk_bd'], pp['sk_wt']), h2pv(pp['sk_salt'] + 'RIGHT', seed.seed, pp['sk_bd'], pp['sk_wt'])
        assert left == sk[0] and right == sk[1] and key_ch * left == vk[0] and key_ch * right == vk[1]
        # make_signature_challenge / sign / verify
        c = h2p(pp['ch_salt'], str(vk) + ', ' + msg, pp['ch_bd'], pp['ch_wt'])
        assert c == lm.make_signature_challenge(pp=pp, otvk=vk, msg=msg)
        sig = left ** c + right
        assert sig == lm.sign(pp=pp, otk=(seed, sk, vk), msg=msg)
        lpp = lm.make_setup_parameters(secpar)
        assert (within(sig, lpp['vf_bd'], lpp['vf_wt']) and key_ch * sig == vk[0] * c + vk[1]) is True
    # aggregation coefficients, aggregate, aggregate_verify
    vks = [k[2] for k in keys]
    sigs = [lm.sign(pp=bpp, otk=k, msg=m) for k, m in zip(keys, msgs)]
    order = sorted(range(2), key=lambda i: str(vks[i]))
    s_keys, s_msgs, s_sigs = [vks[i] for i in order], [msgs[i] for i in order], [sigs[i] for i in order]
    agmsg = str(list(zip(s_keys, s_msgs)))
    coefs = [h2p(bpp['ag_salt'] + str(i), agmsg, bpp['ag_bd'], bpp['ag_wt']) for i in range(2)]
    assert coefs == bk.make_agg_coefs(pp=bpp, otvks=vks, msgs=msgs)
    ag_sig = sum([s ** a for s, a in zip(s_sigs, coefs)])
    assert ag_sig == bk.aggregate(pp=bpp, otvks=vks, msgs=msgs, sigs=sigs)
    challs = [h2p(bpp['ch_salt'], str(k) + ', ' + m, bpp['ch_bd'], bpp['ch_wt']) for k, m in zip(s_keys, s_msgs)]
    total = sum([(k[0] * c + k[1]) * a for a, c, k in zip(coefs, challs, s_keys)])
    assert within(ag_sig, bpp['avf_bd'], bpp['avf_wt'], lower=True) and key_ch * ag_sig == total
    assert bk.aggregate_verify(pp=bpp, otvks=vks, msgs=msgs, ag_sig=ag_sig) is True
    # adaptor: witness / statement, presign, preverify, adapt, verify, extract, witness_verify
    (seed, sk, vk), msg = keys[0], 'Blessed are the cheesemakers.'
    wseed, wit, st = ad.witgen(pp=pp, num_wits_to_gen=1)[0]
    w = h2pv(pp['wit_salt'], wseed.seed, pp['wit_bd'], pp['wit_wt'])
    assert w == wit.key and key_ch * w == st.key
    c = h2p(pp['ch_salt'], str(st) + ', ' + str(vk) + ', ' + msg, pp['ch_bd'], pp['ch_wt'])
    presig = sk[0] ** c + sk[1]
    assert presig == ad.presign(pp=pp, otk=keys[0], msg=msg, st=st)
    assert (within(presig, pp['pvf_bd'], pp['pvf_wt']) and key_ch * presig == vk[0] * c + vk[1]) is True
    assert ad.preverify(pp=pp, otvk=vk, msg=msg, st=st, presig=presig) is True
    full = presig + w
    assert full == ad.adapt(presig=presig, wit=wit)
    assert (within(full, pp['vf_bd'], pp['vf_wt']) and key_ch * full == vk[0] * c + vk[1] + st.key) is True
    assert ad.verify(pp=pp, otvk=vk, msg=msg, st=st, sig=full) is True
    ext = full - presig
    assert ext == ad.extract(pp=pp, presig=presig, sig=full).key == w
    assert (within(ext, pp['ext_wit_bd'], pp['ext_wit_wt']) and key_ch * ext == st.key) is True


def test_aggregation_coefficient_cache_is_keyed_by_content():
    """aggregate followed by aggregate_verify of the same list derives the O(N^2) coefficients once; any change of the
    message content, the index range or the parameters misses."""
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import _ctx, keygen, sign
    pp = bk.set_aggregation_capacity(bk.make_setup_parameters(128), 6)
    keys = keygen(pp=pp, num_keys_to_gen=6)
    msgs = [bin(randbits(32))[2:].zfill(32) for _ in keys]
    sigs = [sign(pp=pp, otk=k, msg=m) for k, m in zip(keys, msgs)]
    vks = [k[2] for k in keys]
    bk.clear_agg_coef_cache()
    eng = _ctx(pp)[0]
    before = eng.launch_count
    ag_sig = bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs)
    mid = eng.launch_count
    assert len(bk._AG_CACHE) == 1
    assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs, ag_sig=ag_sig)
    assert len(bk._AG_CACHE) == 1                                   # hit: same sorted list, same range
    cold = bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs)
    bk.clear_agg_coef_cache()
    assert bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs) == cold == ag_sig          # cached == derived
    assert not bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs[::-1], ag_sig=ag_sig)       # other content: miss, and rejected
    assert len(bk._AG_CACHE) == 2 and mid > before
