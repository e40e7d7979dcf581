"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI with HOST
buffers, against (a) the golden fixtures produced from the reference's own modules
(oracle/gen_golden.py) and (b) the restated CPU oracle on seeded inputs.  Bit-exact everywhere."""
import hashlib
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHIPPED = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20), 256: dict(q=39937, l=23, sk_bd=65, ch_wt=50)}
D = 256


@pytest.fixture(scope='module')
def engines(golden):
    from lattice_cryptography_b200 import Engine
    arrays, _ = golden
    out = {}
    for secpar, p in SHIPPED.items():
        e = Engine(secpar, p['q'], D, p['l'])
        e.set_key_ch(np.ascontiguousarray(arrays[f's{secpar}_key_ch']))
        out[secpar] = e
    yield out
    for e in out.values():
        e.close()


def scheme(secpar):
    from lattice_cryptography_b200 import make_scheme
    p = SHIPPED[secpar]
    return make_scheme(sk_bd=p['sk_bd'], sk_wt=256, ch_bd=1, ch_wt=p['ch_wt'], ag_bd=1, ag_wt=1, wit_bd=1, wit_wt=20)


def negacyclic_mul(a, b, q):
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))
    res = full[:D].copy()
    res[:D - 1] -= full[D:]
    res %= q
    res[res > (q - 1) // 2] -= q
    return res


# ------------------------------------------------------------------------------------------- K1
def test_shake256_matches_hashlib(engines):
    e = engines[128]
    rng = np.random.default_rng(1)
    lens = [0, 1, 7, 8, 9, 134, 135, 136, 137, 271, 272, 273, 1000] + [int(x) for x in rng.integers(0, 700, 300)]
    items = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in lens]
    for out_len in (1, 32, 136, 137, 647):
        got = e.shake256(items, out_len)
        for it, row in zip(items, got):
            assert bytes(row) == hashlib.shake_256(it).digest(out_len)


def test_shake256_empty_batch_and_long_squeeze(engines):
    e = engines[256]
    assert e.shake256([], 10).shape == (0, 10)
    got = e.shake256([b'abc'], 5000)
    assert bytes(got[0]) == hashlib.shake_256(b'abc').digest(5000)


# ------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize('secpar', [128, 256])
def test_sampler_golden_pairs(engines, golden, secpar):
    arrays, meta = golden
    e = engines[secpar]
    m = meta['cases'][str(secpar)]
    p = SHIPPED[secpar]
    # key_ch: bd = q//2 (16-bit limb path), wt = d
    dense, pairs = e.hash2polyvec('KEY_CH_SEED', [meta['key_ch_seed']], p['q'] // 2, D, p['l'], want_pairs=True)
    assert np.array_equal(pairs[0], arrays[f's{secpar}_key_ch_pairs'])
    assert np.array_equal(dense[0], arrays[f's{secpar}_key_ch'])
    # signing key left half of case 0
    dense, pairs = e.hash2polyvec('SK_SALTLEFT', [m['lm'][0]['seed']], p['sk_bd'], D, p['l'], want_pairs=True)
    assert np.array_equal(pairs[0], arrays[f's{secpar}_lm0_skL_pairs'])
    assert np.array_equal(dense[0], arrays[f's{secpar}_lm0_skL'])
    # challenges (ragged messages, wt < d, bd = 1)
    chmsgs = [c['chmsg'] for c in m['lm']]
    dense, pairs = e.hash2polyvec('CH_SALT', chmsgs, 1, p['ch_wt'], 1, want_pairs=True)
    for j in range(len(chmsgs)):
        assert np.array_equal(pairs[j, 0], arrays[f's{secpar}_lm{j}_c_pairs'])
        assert np.array_equal(dense[j, 0], arrays[f's{secpar}_lm{j}_c'])
    # adaptor witnesses
    dense, pairs = e.hash2polyvec('WIT_SALT', [a['wit_seed'] for a in m['adaptor']], 1, 20, p['l'], want_pairs=True)
    for j in range(len(m['adaptor'])):
        assert np.array_equal(pairs[j], arrays[f's{secpar}_ad{j}_wit_pairs'])
        assert np.array_equal(dense[j], arrays[f's{secpar}_ad{j}_wit'])


@pytest.mark.parametrize('secpar,bd,wt,vec_len', [(128, 1, 1, 1), (128, 3, 7, 2), (128, 64, 256, 1), (256, 45, 100, 3),
                                                  (128, 300, 33, 2), (256, 1, 256, 1), (128, 2, 2, 5)])
def test_sampler_vs_oracle_random(engines, secpar, bd, wt, vec_len):
    import lattice_algebra as la
    import schemes
    e = engines[secpar]
    p = SHIPPED[secpar]
    lp = schemes.lattice_parameters(p['q'], D, vec_len)       # vec_len plays `length` for the oracle
    rng = np.random.default_rng(secpar + bd + wt)
    msgs = [''.join(chr(int(c)) for c in rng.integers(32, 127, int(n))) for n in rng.integers(0, 300, 40)]
    dense, pairs = e.hash2polyvec('SALT_X', msgs, bd, wt, vec_len, want_pairs=True)
    bti, btd = la.bits_to_indices(secpar, D, wt), la.bits_to_decode(secpar, bd)
    nb = la.get_gen_bytes_per_poly(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {}, wt, bti, btd)
    for i, msg in enumerate(msgs):
        bits = la.binary_digest(msg, nb * vec_len, 'SALT_X')
        for v in range(vec_len):
            cd = la.decode2polycoefs(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {'bd': bd, 'wt': wt},
                                     bits[v * 8 * nb:(v + 1) * 8 * nb], wt, bti, btd)
            assert [list(map(int, r)) for r in pairs[i, v]] == [[k, c] for k, c in cd.items()]
            want = np.zeros(D, dtype=np.int16)
            for k, c in cd.items():
                want[k] = c
            assert np.array_equal(dense[i, v], want)


# ------------------------------------------------------------------------------------------- K3/K4/K6
@pytest.mark.parametrize('secpar', [128, 256])
def test_ntt_roundtrip_and_product(engines, secpar):
    e = engines[secpar]
    q = SHIPPED[secpar]['q']
    rng = np.random.default_rng(7)
    h = (q - 1) // 2
    a = rng.integers(-h, h + 1, (37, D)).astype(np.int16)
    b = rng.integers(-h, h + 1, (37, D)).astype(np.int16)
    a[0] = 0
    a[1] = h
    b[1] = -h
    a[2, :] = 0
    a[2, 255] = 1          # X^255
    fa = e.ntt_fwd(a)
    assert fa.dtype == np.uint16 and int(fa.max()) < q
    assert np.array_equal(e.ntt_inv(fa), a)
    prod = e.poly_mul(a, b)
    for i in range(a.shape[0]):
        assert np.array_equal(prod[i], negacyclic_mul(a[i], b[i], q))
    # slot p of the NTT form holds a(psi^(2*bitrev8(p)+1)) with psi = least primitive 512th root
    psi = e.root_of_unity
    assert pow(psi, 512, q) == 1 and pow(psi, 256, q) != 1
    x = np.zeros((1, D), dtype=np.int16)
    x[0, 1] = 1            # the polynomial X
    fx = e.ntt_fwd(x)[0]
    for p_ in (0, 1, 2, 77, 255):
        br = int(format(p_, '08b')[::-1], 2)
        assert int(fx[p_]) == pow(psi, 2 * br + 1, q)


@pytest.mark.parametrize('secpar', [128, 256])
def test_reference_ntt_representation_l3(engines, golden, secpar):
    """Parity level L3: the reference's own storage format, Polynomial.ntt_representation (2d-point cyclic
    transform of the zero-padded coefficients, centred, natural order), against the restated lattice_algebra."""
    import lattice_algebra as la
    import schemes
    arrays, _ = golden
    e = engines[secpar]
    p = SHIPPED[secpar]
    lp = schemes.lattice_parameters(p['q'], D, p['l'])
    rng = np.random.default_rng(11)
    h = (p['q'] - 1) // 2
    polys = np.concatenate([arrays[f's{secpar}_lm0_sig'][:3], arrays[f's{secpar}_lm0_vkL'][None],
                            rng.integers(-h, h + 1, (3, D)).astype(np.int16), np.zeros((1, D), np.int16)])
    rep = e.ntt_reference_repr(np.ascontiguousarray(polys))
    assert rep.shape == (polys.shape[0], 2 * D)
    for row, want in zip(polys, rep):
        ref = la.Polynomial(lp=lp, coefs={i: int(v) for i, v in enumerate(row) if v})
        assert want.tolist() == ref.ntt_representation


def test_ntt_any_int16_input(engines):
    e = engines[128]
    q = 11777
    x = np.array([[-32768, 32767] * 128], dtype=np.int16)
    back = e.ntt_inv(e.ntt_fwd(x))[0].astype(np.int64)
    assert np.array_equal((back - x[0]) % q, np.zeros(D, dtype=np.int64))


# ------------------------------------------------------------------------------------------- LM-OTS
@pytest.mark.parametrize('coop', ['1', '0'])          # few streams: cooperative low-latency sampler on / off
@pytest.mark.parametrize('secpar', [128, 256])
def test_lm_golden(engines, golden, monkeypatch, secpar, coop):
    monkeypatch.setenv('LCB_SAMPLER_COOP', coop)
    arrays, meta = golden
    e = engines[secpar]
    m = meta['cases'][str(secpar)]
    sch = scheme(secpar)
    cases = m['lm']
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, [c['seed'] for c in cases])
    for j in range(len(cases)):
        pre = f's{secpar}_lm{j}'
        assert np.array_equal(sk_coef[j, 0], arrays[pre + '_skL'])
        assert np.array_equal(sk_coef[j, 1], arrays[pre + '_skR'])
        assert np.array_equal(vk_coef[j, 0], arrays[pre + '_vkL'])
        assert np.array_equal(vk_coef[j, 1], arrays[pre + '_vkR'])
    assert np.array_equal(e.ntt_inv(vk_ntt), vk_coef)
    assert np.array_equal(e.ntt_inv(sk_ntt), sk_coef)
    chmsgs = [c['chmsg'] for c in cases]
    sig = e.lm_sign(sch, sk_ntt, chmsgs)
    for j in range(len(cases)):
        assert np.array_equal(sig[j], arrays[f's{secpar}_lm{j}_sig'])
    vf_bd, vf_wt = m['vf_bd'], m['vf_wt']
    assert e.lm_verify(sch, vk_ntt, chmsgs, sig, vf_bd, vf_wt).tolist() == [int(c['verdict']) for c in cases]
    bad = [c['chmsg'] + '!' for c in cases]
    assert e.lm_verify(sch, vk_ntt, bad, sig, vf_bd, vf_wt).tolist() == [int(c['verdict_bad_msg']) for c in cases]
    t1 = np.stack([arrays[f's{secpar}_lm{j}_sig_t1'] for j in range(len(cases))])
    t2 = np.stack([arrays[f's{secpar}_lm{j}_sig_t2'] for j in range(len(cases))])
    assert e.lm_verify(sch, vk_ntt, chmsgs, t1, vf_bd, vf_wt).tolist() == [int(c['verdict_t1']) for c in cases]
    assert e.lm_verify(sch, vk_ntt, chmsgs, t2, vf_bd, vf_wt).tolist() == [int(c['verdict_t2']) for c in cases]
    # boundary of the norm test: a bound equal to the actual max norm passes, one below fails
    n0 = int(np.abs(sig[0]).max())
    assert e.lm_verify(sch, vk_ntt[:1], chmsgs[:1], sig[:1], n0, vf_wt).tolist() == [1]
    assert e.lm_verify(sch, vk_ntt[:1], chmsgs[:1], sig[:1], n0 - 1, vf_wt).tolist() == [0]
    # weight test: every signature polynomial here is dense, so wt = 255 may reject
    w0 = int((sig[0] != 0).sum(axis=1).max())
    assert e.lm_verify(sch, vk_ntt[:1], chmsgs[:1], sig[:1], vf_bd, w0).tolist() == [1]
    assert e.lm_verify(sch, vk_ntt[:1], chmsgs[:1], sig[:1], vf_bd, w0 - 1).tolist() == [0]


@pytest.mark.parametrize('secpar,n', [(128, 300), (256, 131)])
def test_lm_batch_vs_oracle(engines, golden, secpar, n):
    """Seeded batch with ragged messages: every sk/vk/sig coefficient against the CPU oracle on a
    subsample, every verdict against the construction rule (every 5th triple tampered)."""
    import schemes
    arrays, meta = golden
    e = engines[secpar]
    sch = scheme(secpar)
    m = meta['cases'][str(secpar)]
    rng = np.random.default_rng(20260101 + secpar)
    seeds = [''.join(rng.choice(['0', '1'], secpar)) for _ in range(n)]
    chmsgs = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{i:010x}>, ' +
              ''.join(rng.choice(['0', '1'], int(rng.integers(0, 2 * secpar)))) for i in range(n)]
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
    sig = e.lm_sign(sch, sk_ntt, chmsgs)
    tampered = sig.copy()
    expect = np.ones(n, dtype=np.uint8)
    for i in range(0, n, 5):
        kind = (i // 5) % 3
        if kind == 0:
            tampered[i, i % tampered.shape[1], (3 * i) % D] += 1
        elif kind == 1:
            tampered[i, (i + 1) % tampered.shape[1], (7 * i) % D] = m['vf_bd'] + 1
        expect[i] = 0
    msgs2 = list(chmsgs)
    for i in range(0, n, 5):
        if (i // 5) % 3 == 2:
            msgs2[i] = chmsgs[i] + '0'
    got = e.lm_verify(sch, vk_ntt, msgs2, tampered, m['vf_bd'], m['vf_wt'])
    assert np.array_equal(got, expect)
    assert e.lm_verify(sch, vk_ntt, chmsgs, sig, m['vf_bd'], m['vf_wt']).all()
    # oracle subsample
    key_ch = schemes.vec_from_dense(schemes.lattice_parameters(SHIPPED[secpar]['q'], D, SHIPPED[secpar]['l']),
                                    arrays[f's{secpar}_key_ch'].tolist())
    pp = schemes.make_lm_parameters(secpar, key_ch)
    for i in (0, n // 2, n - 1):
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, seeds[i])
        assert schemes.dense_of_vec(skl) == sk_coef[i, 0].tolist()
        assert schemes.dense_of_vec(skr) == sk_coef[i, 1].tolist()
        assert schemes.dense_of_poly(vkl) == vk_coef[i, 0].tolist()
        assert schemes.dense_of_poly(vkr) == vk_coef[i, 1].tolist()
        osig = schemes.lm_sign(pp, skl, skr, chmsgs[i])
        assert schemes.dense_of_vec(osig) == sig[i].tolist()
        assert schemes.lm_verify(pp, vkl, vkr, msgs2[i], schemes.vec_from_dense(pp['lp'], tampered[i].tolist())) == \
            bool(expect[i])


def test_device_buffers_and_empty_batches(engines, golden):
    import torch
    arrays, meta = golden
    e = engines[128]
    sch = scheme(128)
    cases = meta['cases']['128']['lm']
    seeds = [c['seed'] for c in cases]
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds, device=True)
    assert sk_ntt.is_cuda
    from lattice_cryptography_b200 import ragged
    blob, off = ragged([c['chmsg'] for c in cases])
    dmsg = (torch.from_numpy(blob.copy()).cuda(), torch.from_numpy(off).cuda())
    sig = e.lm_sign(sch, sk_ntt, dmsg, device=True)
    verdict = e.lm_verify(sch, vk_ntt, dmsg, sig, 945, 256, device=True)
    e.synchronize()
    assert verdict.cpu().tolist() == [1] * len(cases)
    assert np.array_equal(sig.cpu().numpy()[0], arrays['s128_lm0_sig'])
    # empty batches are accepted everywhere
    assert e.lm_verify(sch, np.zeros((0, 2, D), np.uint16), [], np.zeros((0, 13, D), np.int16), 945, 256).shape == (0,)
    assert e.lm_sign(sch, np.zeros((0, 2, 13, D), np.uint16), []).shape == (0, 13, D)


# ------------------------------------------------------------------------------------------- BKLM
@pytest.mark.parametrize('secpar', [128, 256])
def test_bklm_golden(engines, golden, secpar):
    arrays, meta = golden
    e = engines[secpar]
    sch = scheme(secpar)
    m = meta['cases'][str(secpar)]
    seeds = [c['seed'] for c in m['lm']]
    _, sk_ntt, vk_ntt, _ = e.lm_keygen(sch, seeds)
    for case in m['bklm']:
        cap = case['cap']
        pre = f's{secpar}_bk{cap}'
        sigs = e.lm_sign(sch, sk_ntt[:cap], case['chmsgs'])
        assert np.array_equal(sigs, arrays[pre + '_sigs'])
        order = case['sorted_order']
        ag = e.agg_coefs(sch, case['agmsg'], 0, cap)
        want = arrays[pre + '_ag_coefs']
        for i in range(cap):
            k, s = int(ag[i, 0, 0]), int(ag[i, 0, 1])
            dense = np.zeros(D, dtype=np.int16)
            dense[k] = s
            assert np.array_equal(dense, want[i])
        # split the sorted list into two shards, as two ranks would
        srt_sigs = np.ascontiguousarray(sigs[order])
        cut = cap // 2
        ag_hi = e.agg_coefs(sch, case['agmsg'], cut, cap - cut)
        assert np.array_equal(ag_hi, ag[cut:])
        part = e.aggregate_partial(sch, srt_sigs[:cut], np.ascontiguousarray(ag[:cut])).astype(np.int64) + \
            e.aggregate_partial(sch, srt_sigs[cut:], ag_hi)
        ag_sig = e.aggregate_finish(part.astype(np.int32))
        assert np.array_equal(ag_sig, arrays[pre + '_ag_sig'])
        srt_vk = np.ascontiguousarray(vk_ntt[:cap][order])
        srt_ch = [case['chmsgs'][i] for i in order]
        vp = e.aggverify_partial(sch, srt_vk[:cut], srt_ch[:cut], np.ascontiguousarray(ag[:cut])).astype(np.int64) + \
            e.aggverify_partial(sch, srt_vk[cut:], srt_ch[cut:], ag_hi)
        vp = vp.astype(np.int32)
        assert e.aggverify_finish(vp, ag_sig, cap, cap, case['avf_bd'], case['avf_wt']) == case['verdict']
        bad = ag_sig.copy()
        bad[0, 0] += 1
        assert e.aggverify_finish(vp, bad, cap, cap, case['avf_bd'], case['avf_wt']) == case['verdict_tampered']
        # bound / cap failures (bklm_one_time_agg_sigs.py:103-105)
        assert not e.aggverify_finish(vp, ag_sig, cap, cap - 1, case['avf_bd'], case['avf_wt'])
        assert not e.aggverify_finish(vp, ag_sig, cap, cap, int(np.abs(ag_sig).max()) - 1, case['avf_wt'])
        assert not e.aggverify_finish(vp, np.zeros_like(ag_sig), cap, cap, case['avf_bd'], case['avf_wt'])


@pytest.mark.parametrize('lanes', [1, 2])
@pytest.mark.parametrize('n,first,msg_len', [(1, 0, 0), (130, 0, 5), (257, 95, 1000), (1500, 9990, 33000), (64, 99999990, 271),
                                             (40, 0, 7), (40, 0, 8), (33, 90, 121), (33, 90, 128), (33, 90, 129),
                                             (19, 9_999_999_990, 4000), (70, 999_999_999_999_999_980, 300)])
def test_agg_coefs_vs_hashlib(engines, monkeypatch, n, first, msg_len, lanes):
    """make_agg_coefs (bklm_one_time_agg_sigs.py:78-81) for many indices over one shared message:
    ag_i = +-X^k with k = first digest byte, sign = next bit of SHAKE256('AG_SALT' + str(i) + msg).
    Covers every salt-length class (1..18 digits, i.e. every byte phase of the message against the 64-bit words of
    the sponge), digit-count changes inside a warp (streams of one warp then need different numbers of rate blocks),
    messages shorter than a word / a block, unaligned messages and sharded index ranges - for both kernels: one
    thread per sponge (k_agg_coefs) and two lanes per sponge over the pre-split message (k_agg_coefs_il)."""
    monkeypatch.setenv('LCB_AGG_LANES', str(lanes))
    e = engines[128]
    sch = scheme(128)
    rng = np.random.default_rng(n + msg_len)
    msg = bytes(rng.integers(32, 127, msg_len, dtype=np.uint8))
    blob = np.frombuffer(b'xyz' + msg, dtype=np.uint8)[3:]          # deliberately misaligned host view
    got = e.agg_coefs(sch, np.ascontiguousarray(blob) if msg_len == 0 else blob, first, n)
    for i in range(n):
        dg = hashlib.shake_256(b'AG_SALT' + str(first + i).encode() + msg).digest(2)
        assert (int(got[i, 0, 0]), int(got[i, 0, 1])) == (dg[0], 1 if dg[1] & 0x80 else -1), (i, first + i)


# ------------------------------------------------------------------------------------------- adaptor
@pytest.mark.parametrize('coop', ['1', '0'])
@pytest.mark.parametrize('secpar', [128, 256])
def test_adaptor_golden(engines, golden, monkeypatch, secpar, coop):
    monkeypatch.setenv('LCB_SAMPLER_COOP', coop)
    arrays, meta = golden
    e = engines[secpar]
    sch = scheme(secpar)
    m = meta['cases'][str(secpar)]
    ap = m['adaptor_params']
    cases = m['adaptor']
    seeds = [m['lm'][c['key_index']]['seed'] for c in cases]
    _, sk_ntt, vk_ntt, _ = e.lm_keygen(sch, seeds)
    wit, st_ntt, st_coef = e.witgen(sch, [c['wit_seed'] for c in cases])
    chmsgs = [c['chmsg'] for c in cases]
    presig = e.lm_sign(sch, sk_ntt, chmsgs)
    sig = e.vec_add(presig, wit)
    ext = e.vec_sub(sig, presig)
    for j in range(len(cases)):
        pre = f's{secpar}_ad{j}'
        assert np.array_equal(wit[j], arrays[pre + '_wit'])
        assert np.array_equal(st_coef[j], arrays[pre + '_st'])
        assert np.array_equal(presig[j], arrays[pre + '_presig'])
        assert np.array_equal(sig[j], arrays[pre + '_sig'])
        assert np.array_equal(ext[j], arrays[pre + '_ext'])
    assert np.array_equal(e.ntt_inv(st_ntt), st_coef)
    pv = e.lm_verify(sch, vk_ntt, chmsgs, presig, ap['pvf_bd'], ap['pvf_wt'])
    assert pv.tolist() == [int(c['preverify']) for c in cases]
    vv = e.lm_verify(sch, vk_ntt, chmsgs, sig, ap['vf_bd'], ap['vf_wt'], st_ntt=st_ntt)
    assert vv.tolist() == [int(c['verify']) for c in cases]
    wv = e.witness_verify(ext, st_ntt, ap['ext_wit_bd'], ap['ext_wit_wt'])
    assert wv.tolist() == [int(c['witness_verify']) for c in cases]
    pa = e.lm_verify(sch, vk_ntt, chmsgs, sig, ap['pvf_bd'], ap['pvf_wt'])
    assert pa.tolist() == [int(c['preverify_of_adapted']) for c in cases]
    # a wrong witness does not verify
    assert e.witness_verify(np.ascontiguousarray(ext[::-1]), st_ntt, ap['ext_wit_bd'], ap['ext_wit_wt']).tolist() == [0, 0]


# ------------------------------------------------------------------------------------------- big batches vs the C oracle
@pytest.mark.parametrize('secpar,n', [(128, 2048), (256, 600)])
def test_lm_big_batch_vs_c_oracle(engines, golden, secpar, n):
    """Every verdict of a few thousand triples (a third tampered in assorted ways, ragged messages)
    and every key / signature coefficient of a subsample against oracle/lcb_oracle.c, whose products
    are schoolbook convolutions and whose decoder walks the digest bit by bit."""
    import c_oracle
    arrays, meta = golden
    e = engines[secpar]
    sch = scheme(secpar)
    s = SHIPPED[secpar]
    m = meta['cases'][str(secpar)]
    p = c_oracle.params(secpar, s['q'], s['l'], s['sk_bd'], s['ch_wt'])
    key_ch = np.ascontiguousarray(arrays[f's{secpar}_key_ch'])
    rng = np.random.default_rng(99 + secpar)
    seeds = [''.join(rng.choice(['0', '1'], secpar + int(rng.integers(0, 9)))) for _ in range(n)]
    chmsgs = [bytes(rng.integers(1, 256, int(rng.integers(0, 400)), dtype=np.uint8)) for _ in range(n)]
    sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
    sig = e.lm_sign(sch, sk_ntt, chmsgs)
    bad = sig.copy()
    msgs2 = list(chmsgs)
    for i in range(n):
        kind = int(rng.integers(0, 9))
        a, b = int(rng.integers(0, s['l'])), int(rng.integers(0, 256))
        if kind == 0:
            bad[i, a, b] += int(rng.choice([-1, 1]))
        elif kind == 1:
            bad[i, a, b] = int(rng.choice([-1, 1])) * (m['vf_bd'] + int(rng.integers(0, 3)))   # at / over the bound
        elif kind == 2:
            msgs2[i] = chmsgs[i] + b'\x01'
        elif kind == 3:
            bad[i] = sig[(i + 1) % n]
    from lattice_cryptography_b200 import ragged
    blob, off = ragged(msgs2)
    got = e.lm_verify(sch, vk_ntt, (blob, off), bad, m['vf_bd'], m['vf_wt'])
    want = c_oracle.lm_verify_batch(p, key_ch, vk_coef, blob, off, bad, m['vf_bd'], m['vf_wt'])
    assert np.array_equal(got, want)
    assert 0.5 < want.mean() < 0.8          # both verdicts are well represented
    for i in range(0, n, max(1, n // 24)):
        skl, skr, vkl, vkr = c_oracle.lm_keygen(p, key_ch, seeds[i].encode())
        assert np.array_equal(skl, sk_coef[i, 0]) and np.array_equal(skr, sk_coef[i, 1])
        assert np.array_equal(vkl, vk_coef[i, 0]) and np.array_equal(vkr, vk_coef[i, 1])
        assert np.array_equal(c_oracle.lm_sign(p, skl, skr, chmsgs[i]), sig[i])


# ------------------------------------------------------------------------------------------- error contract of the ABI
def test_abi_error_paths():
    from lattice_cryptography_b200 import Engine, LcbError, _ffi, make_scheme
    e = Engine(128, 11777, D, 13)
    sch = scheme(128)
    try:
        with pytest.raises(LcbError) as err:            # row-vector products need key_ch first
            e.lm_verify(sch, np.zeros((1, 2, D), np.uint16), ['m'], np.zeros((1, 13, D), np.int16), 945, 256)
        assert err.value.status == _ffi.LCB_ERR_NO_KEY_CH
        with pytest.raises(LcbError) as err:
            e.lm_keygen(sch, ['0' * 128])
        assert err.value.status == _ffi.LCB_ERR_NO_KEY_CH
        # samplers and transforms do not need key_ch
        dense, _ = e.hash2polyvec('S', ['x'], 5, 7, 2)
        assert (dense != 0).sum(axis=-1).tolist() == [[7, 7]] and int(np.abs(dense).max()) <= 5
        for bad in (dict(bd=0, wt=1), dict(bd=1, wt=0), dict(bd=1, wt=257), dict(bd=40000, wt=1)):
            with pytest.raises(LcbError) as err:
                e.hash2polyvec('S', ['x'], bad['bd'], bad['wt'], 1)
            assert err.value.status == _ffi.LCB_ERR_INVALID
        with pytest.raises(LcbError) as err:            # aggregation weights beyond the degree are rejected
            e.agg_coefs(make_scheme(ag_wt=257), 'msg', 0, 4)
        assert err.value.status == _ffi.LCB_ERR_INVALID
        with pytest.raises(ValueError):                 # non-contiguous host buffers are refused by the shim
            e.ntt_fwd(np.zeros((4, 2 * D), np.int16)[:, ::2])
    finally:
        e.close()


@pytest.mark.parametrize('secpar', [1, 7, 24, 64, 200, 384, 512])
def test_sampler_field_widths_vs_c_oracle(secpar):
    """The decoder's fields are (8 + secpar) and (ceil(log2 bd) + 1 + secpar) bits wide: exercise widths from
    9 to 528 bits (window refills at every alignment), all three modulus paths (bd = 1, bd <= 256, bd > 256)
    and weights from 1 to 256, against the bit-by-bit C oracle."""
    import c_oracle
    from lattice_cryptography_b200 import Engine
    e = Engine(secpar, 11777, D, 2)
    try:
        rng = np.random.default_rng(secpar)
        cases = [(1, 1), (1, 256), (2, 3), (45, 256), (255, 17), (256, 64), (257, 5), (5888, 256), (5000, 31)]
        for bd, wt in cases:
            msgs = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in (0, 1, 135, 136, 137, 300)]
            dense, pairs = e.hash2polyvec('S' * (secpar % 9), msgs, bd, wt, 2, want_pairs=True)
            for i, msg in enumerate(msgs):
                od, op = c_oracle.hash2polyvec(secpar, D, 'S' * (secpar % 9), msg, bd, wt, 2)
                assert np.array_equal(pairs[i], op), (secpar, bd, wt, i)
                assert np.array_equal(dense[i], od), (secpar, bd, wt, i)
    finally:
        e.close()


# ------------------------------------------------------------------------------------------- absorb edge cases
@pytest.mark.parametrize('salt', ['S', 'SALT', 'SALT_XY', 'SALT_12CHARS'])
def test_absorb_block_boundaries_vs_c_oracle(engines, salt):
    """Every message length around the 136-byte rate boundaries (salt + message = 135..138, 271..274, 407..409,
    plus the empty message), at every byte alignment of the message inside the blob: the batched block loader of
    the sampler (sampler_device.cuh, InputView::load_block) against oracle/lcb_oracle.c's byte-wise sponge."""
    import c_oracle
    e = engines[128]
    rng = np.random.default_rng(len(salt))
    targets = [0, 1, 2, 3, 4, 5] + [t - len(salt) + d for t in (136, 272, 408) for d in (-3, -2, -1, 0, 1, 2, 3)]
    lens = sorted({n for n in targets if n >= 0})
    msgs = []
    for n in lens:
        for _ in range(4):            # four copies at different blob offsets (the pad strings shift the alignment)
            msgs.append(bytes(rng.integers(0, 256, n, dtype=np.uint8)))
            msgs.append(bytes(rng.integers(0, 256, int(rng.integers(0, 4)), dtype=np.uint8)))
    dense, pairs = e.hash2polyvec(salt, msgs, 1, 20, 1, want_pairs=True)
    for i, m in enumerate(msgs):
        want, want_pairs = c_oracle.hash2polyvec(128, D, salt, m, 1, 20, 1)
        assert np.array_equal(dense[i], want) and np.array_equal(pairs[i], want_pairs), (salt, len(m), i)


def test_keygen_seed_lengths_across_block_boundary(engines, golden):
    """Paired key-generation launch: SK_SALTLEFT (11 bytes) and SK_SALTRIGHT (12 bytes) put the two halves of one
    key on different sides of a rate boundary for seeds of 124 / 260 bytes - lanes of one warp then absorb a
    different number of blocks."""
    import c_oracle
    arrays, _ = golden
    for secpar, base in ((128, 120), (256, 256)):
        s = SHIPPED[secpar]
        e = engines[secpar]
        p = c_oracle.params(secpar, s['q'], s['l'], s['sk_bd'], s['ch_wt'])
        key_ch = np.ascontiguousarray(arrays[f's{secpar}_key_ch'])
        rng = np.random.default_rng(secpar)
        seeds = [''.join(rng.choice(['0', '1'], base + k)) for k in range(0, 10)]
        sk_coef, _, _, vk_coef = e.lm_keygen(scheme(secpar), seeds)
        for i in (0, 3, 4, 5, 9):
            skl, skr, vkl, vkr = c_oracle.lm_keygen(p, key_ch, seeds[i].encode())
            assert np.array_equal(skl, sk_coef[i, 0]) and np.array_equal(skr, sk_coef[i, 1]), (secpar, len(seeds[i]))
            assert np.array_equal(vkl, vk_coef[i, 0]) and np.array_equal(vkr, vk_coef[i, 1])


# ------------------------------------------------------------------------------------------- other moduli
def _ntt_primes():
    ps = [q for q in range(513, 65536, 512) if all(q % f for f in range(2, int(q ** 0.5) + 1))]
    return [ps[0], ps[len(ps) // 2], ps[-1]]        # smallest, a middle one, the largest below 2^16


@pytest.mark.parametrize('q', _ntt_primes())
@pytest.mark.parametrize('l', [1, 2, 7])
def test_other_ntt_friendly_moduli_vs_c_oracle(q, l):
    """lcb_ctx_create accepts every prime q < 2^16 with q = 1 mod 512 and 1 <= l <= 64: the FP32-assisted
    butterflies, the lazy bounds and the tables are generic in q, and l = 1 drives the 2-deep staging pipelines
    across item boundaries on every step.  keygen -> sign -> verify (with tampering) against lcb_oracle.c."""
    import c_oracle
    from lattice_cryptography_b200 import Engine, make_scheme, ragged
    secpar, sk_bd, ch_wt, n = 128, 5, 20, 300
    vf_bd = min(q // 2, sk_bd * (1 + ch_wt))
    rng = np.random.default_rng(q + l)
    key_ch = rng.integers(-(q // 2), q // 2 + 1, size=(l, D)).astype(np.int16)
    e = Engine(secpar, q, D, l)
    try:
        e.set_key_ch(key_ch)
        sch = make_scheme(sk_bd=sk_bd, sk_wt=256, ch_bd=1, ch_wt=ch_wt)
        p = c_oracle.params(secpar, q, l, sk_bd, ch_wt)
        seeds = [''.join(rng.choice(['0', '1'], secpar)) for _ in range(n)]
        chmsgs = [bytes(rng.integers(1, 256, int(rng.integers(0, 200)), dtype=np.uint8)) for _ in range(n)]
        sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
        sig = e.lm_sign(sch, sk_ntt, chmsgs)
        for i in (0, 1, n - 1):
            skl, skr, vkl, vkr = c_oracle.lm_keygen(p, key_ch, seeds[i].encode())
            assert np.array_equal(skl, sk_coef[i, 0]) and np.array_equal(vkl, vk_coef[i, 0]) and np.array_equal(vkr, vk_coef[i, 1])
            assert np.array_equal(c_oracle.lm_sign(p, skl, skr, chmsgs[i]), sig[i])
        bad = sig.copy()
        for i in range(0, n, 3):
            bad[i, int(rng.integers(0, l)), int(rng.integers(0, D))] += int(rng.choice([-1, 1]))
        blob, off = ragged(chmsgs)
        got = e.lm_verify(sch, vk_ntt, (blob, off), bad, vf_bd, 256)
        want = c_oracle.lm_verify_batch(p, key_ch, vk_coef, blob, off, bad, vf_bd, 256)
        assert np.array_equal(got, want) and want.sum() == n - len(range(0, n, 3))
        # round trip and product through the unit-surface kernels
        a = rng.integers(-(q // 2), q // 2 + 1, size=(5, D)).astype(np.int16)
        b = rng.integers(-(q // 2), q // 2 + 1, size=(5, D)).astype(np.int16)
        assert np.array_equal(e.ntt_inv(e.ntt_fwd(a)), a)
        assert np.array_equal(e.poly_mul(a, b), np.stack([negacyclic_mul(x, y, q) for x, y in zip(a, b)]).astype(np.int16))
    finally:
        e.close()


@pytest.mark.parametrize('secpar', [128, 256])
@pytest.mark.parametrize('n', [1, 2, 37, 74, 75])
def test_cooperative_sampler_equals_one_thread_per_stream(engines, monkeypatch, secpar, n):
    """k_sampler_coop (one block per stream: a lane-pair sponge feeding one decoder warp per polynomial) against
    k_sampler on the same seeds: keys (paired mode: 2 n streams, up to one block per SM), witnesses and a generic
    hash2polynomialvector call, including seeds whose salt || seed crosses a rate block."""
    e = engines[secpar]
    sch = scheme(secpar)
    seeds = [bin(3 ** (j + 5))[2:].zfill(secpar)[-secpar:] + '01' * (j % 4) for j in range(n)]
    out = {}
    for coop in ('0', '1'):
        monkeypatch.setenv('LCB_SAMPLER_COOP', coop)
        keys = e.lm_keygen(sch, seeds)
        wit = e.witgen(sch, seeds)
        dense, pairs = e.hash2polyvec('COOP', seeds, 45, 256, 5, want_pairs=True)
        sparse, sp = e.hash2polyvec('COOP', seeds, 7, 33, 3, want_pairs=True)
        out[coop] = keys + wit + (dense, pairs, sparse, sp)
    for a, b in zip(out['0'], out['1']):
        assert np.array_equal(a, b)
