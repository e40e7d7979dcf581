"""CPU tests of the oracle itself: the restated scheme layer (oracle/schemes.py) must reproduce the
golden fixtures, which were produced by the REFERENCE'S OWN modules (oracle/gen_golden.py), from
nothing but the recorded seeds and hash-input strings.  Where /root/reference is present the two
are also run side by side.  SHAKE256 is pinned by hashlib."""
import hashlib

import numpy as np
import pytest

import lattice_algebra as la          # the restatement (tests/conftest.py puts oracle/ on sys.path)
import ref_loader
import schemes


def _pp(secpar, arrays, kind='lm', ag_cap=None):
    s = schemes.SHIPPED[secpar]
    lp = schemes.lattice_parameters(s['modulus'], s['degree'], s['length'])
    key_ch = schemes.vec_from_dense(lp, arrays[f's{secpar}_key_ch'].tolist())
    if kind == 'lm':
        return schemes.make_lm_parameters(secpar, key_ch)
    if kind == 'bklm':
        return schemes.make_bklm_parameters(secpar, key_ch, ag_cap)
    return schemes.make_adaptor_parameters(secpar, key_ch)


def test_parameter_tables():
    for secpar, (q, rou, vf_bd) in {128: (11777, 24, 945), 256: (39937, 55, 3315)}.items():
        s = schemes.SHIPPED[secpar]
        lp = schemes.lattice_parameters(s['modulus'], s['degree'], s['length'])
        assert (lp.modulus, lp.rou, lp.n, lp.halfmod) == (q, rou, 512, q // 2)
        assert pow(lp.rou, 512, q) == 1 and pow(lp.rou, 256, q) == q - 1
    assert la.bits_to_indices(128, 256, 256) + 256 * la.bits_to_decode(128, 45) == 69248
    assert la.bits_to_indices(256, 256, 50) + 50 * la.bits_to_decode(256, 1) == 25794
    assert la.is_bitstring('') and la.is_bitstring('0110') and not la.is_bitstring('012') and not la.is_bitstring(5)
    assert la.is_ntt_friendly_prime(193, 32) and not la.is_ntt_friendly_prime(197, 32)


def test_key_ch_is_the_seeded_row(golden):
    arrays, meta = golden
    for secpar in (128,):
        kc = schemes.key_ch_from_seed(secpar, meta['key_ch_seed'])
        assert np.array_equal(np.array(schemes.dense_of_vec(kc), dtype=np.int16), arrays[f's{secpar}_key_ch'])


@pytest.mark.parametrize('secpar', [128, 256])
def test_lm_golden(golden, secpar):
    arrays, meta = golden
    pp = _pp(secpar, arrays)
    m = meta['cases'][str(secpar)]
    assert (pp['vf_bd'], pp['vf_wt']) == (m['vf_bd'], m['vf_wt'])
    for j, case in enumerate(m['lm'][:2 if secpar == 256 else 4]):
        pre = f's{secpar}_lm{j}'
        digest = hashlib.shake_256(('SK_SALTLEFT' + case['seed']).encode()).digest(4096)
        assert digest[:64].hex() == case['skL_digest_first64']
        assert hashlib.sha256(digest).hexdigest() == case['skL_digest4096_sha256']
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, case['seed'])
        assert np.array_equal(np.array(schemes.dense_of_vec(skl)), arrays[pre + '_skL'])
        assert np.array_equal(np.array(schemes.dense_of_vec(skr)), arrays[pre + '_skR'])
        assert np.array_equal(np.array(schemes.dense_of_poly(vkl)), arrays[pre + '_vkL'])
        assert np.array_equal(np.array(schemes.dense_of_poly(vkr)), arrays[pre + '_vkR'])
        c = schemes.challenge(pp, case['chmsg'])
        assert np.array_equal(np.array(schemes.dense_of_poly(c)), arrays[pre + '_c'])
        sig = schemes.lm_sign(pp, skl, skr, case['chmsg'])
        assert np.array_equal(np.array(schemes.dense_of_vec(sig)), arrays[pre + '_sig'])
        assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'], sig) == case['verdict']
        assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'] + '!', sig) == case['verdict_bad_msg']
        for tag in ('t1', 't2'):
            bad = schemes.vec_from_dense(pp['lp'], arrays[f'{pre}_sig_{tag}'].tolist())
            assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'], bad) == case[f'verdict_{tag}']


def test_bklm_golden(golden):
    arrays, meta = golden
    secpar = 128
    m = meta['cases'][str(secpar)]
    for case in m['bklm']:
        cap = case['cap']
        pp = _pp(secpar, arrays, 'bklm', cap)
        assert (pp['avf_bd'], pp['avf_wt']) == (case['avf_bd'], case['avf_wt'])
        lp = pp['lp']
        pre = f's{secpar}_bk{cap}'
        order = case['sorted_order']
        sigs = [schemes.vec_from_dense(lp, arrays[pre + '_sigs'][i].tolist()) for i in order]
        coefs = schemes.agg_coefs(pp, case['agmsg'], cap)
        assert np.array_equal(np.array([schemes.dense_of_poly(c) for c in coefs]), arrays[pre + '_ag_coefs'])
        ag_sig = schemes.aggregate(pp, sigs, case['agmsg'])
        assert np.array_equal(np.array(schemes.dense_of_vec(ag_sig)), arrays[pre + '_ag_sig'])
        vks = [(schemes.poly_from_dense(lp, arrays[f's{secpar}_lm{i}_vkL'].tolist()),
                schemes.poly_from_dense(lp, arrays[f's{secpar}_lm{i}_vkR'].tolist())) for i in order]
        chm = [case['chmsgs'][i] for i in order]
        assert schemes.aggregate_verify(pp, vks, chm, case['agmsg'], ag_sig) == case['verdict']
        bad = arrays[pre + '_ag_sig'].copy()
        bad[0, 0] += 1
        assert schemes.aggregate_verify(pp, vks, chm, case['agmsg'], schemes.vec_from_dense(lp, bad.tolist())) == \
            case['verdict_tampered']


def test_adaptor_golden(golden):
    arrays, meta = golden
    secpar = 128
    m = meta['cases'][str(secpar)]
    pp = _pp(secpar, arrays, 'adaptor')
    for k, v in m['adaptor_params'].items():
        assert pp[k] == v
    lp = pp['lp']
    for j, case in enumerate(m['adaptor']):
        pre = f's{secpar}_ad{j}'
        wit, st = schemes.witgen_one(pp, case['wit_seed'])
        assert np.array_equal(np.array(schemes.dense_of_vec(wit)), arrays[pre + '_wit'])
        assert np.array_equal(np.array(schemes.dense_of_poly(st)), arrays[pre + '_st'])
        i = case['key_index']
        skl, skr, vkl, vkr = schemes.lm_keygen_one(pp, m['lm'][i]['seed'])
        presig = schemes.lm_sign(pp, skl, skr, case['chmsg'])
        assert np.array_equal(np.array(schemes.dense_of_vec(presig)), arrays[pre + '_presig'])
        assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'], presig, None, 'pvf_bd', 'pvf_wt') == case['preverify']
        sig = schemes.adapt(presig, wit)
        assert np.array_equal(np.array(schemes.dense_of_vec(sig)), arrays[pre + '_sig'])
        assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'], sig, st) == case['verify']
        assert schemes.lm_verify(pp, vkl, vkr, case['chmsg'], sig, None, 'pvf_bd', 'pvf_wt') == case['preverify_of_adapted']
        ext = schemes.extract(presig, sig)
        assert np.array_equal(np.array(schemes.dense_of_vec(ext)), arrays[pre + '_ext'])
        assert schemes.witness_verify(pp, ext, st) == case['witness_verify']


@pytest.mark.skipif(not ref_loader.reference_available(), reason='reference checkout not present on this box')
def test_oracle_equals_reference_modules_live():
    """The restated scheme layer against the reference's own modules in the same process."""
    otk, lm, bklm, ad = ref_loader.load_reference()
    secpar = 128
    pp = lm.make_setup_parameters(secpar)
    lp = pp['scheme_parameters'].lp
    key_ch = schemes.key_ch_from_seed(secpar, 'live cross-check')
    pp['scheme_parameters'].key_ch = key_ch
    opp = schemes.make_lm_parameters(secpar, key_ch)
    assert {k: opp[k] for k in ('sk_bd', 'sk_wt', 'ch_bd', 'ch_wt', 'vf_bd', 'vf_wt')} == \
        {k: pp[k] for k in ('sk_bd', 'sk_wt', 'ch_bd', 'ch_wt', 'vf_bd', 'vf_wt')}
    seed = bin(424242)[2:].zfill(secpar)
    key = lm.keygen_core(pp=pp, num_keys_to_gen=1, seeds=[otk.SecretSeed(secpar=secpar, lp=lp, seed=seed)])[0]
    skl, skr, vkl, vkr = schemes.lm_keygen_one(opp, seed)
    assert key[1][0] == skl and key[1][1] == skr and key[2][0] == vkl and key[2][1] == vkr
    msg = '0110' * 32
    sig = lm.sign(pp=pp, otk=key, msg=msg)
    chmsg = str(key[2]) + ', ' + msg
    assert sig == schemes.lm_sign(opp, skl, skr, chmsg)
    assert lm.verify(pp=pp, otvk=key[2], msg=msg, sig=sig) is True
    assert schemes.lm_verify(opp, vkl, vkr, chmsg, sig) is True
    assert schemes.lm_verify(opp, vkl, vkr, chmsg + '0', sig) is False
    # parameter tables of the other two schemes
    bpp, app = bklm.make_setup_parameters(secpar), ad.make_setup_parameters(secpar)
    obp, oap = schemes.make_bklm_parameters(secpar, key_ch), schemes.make_adaptor_parameters(secpar, key_ch)
    for k in ('ag_cap', 'ag_bd', 'ag_wt', 'avf_bd', 'avf_wt', 'ag_salt'):
        assert bpp[k] == obp[k]
    for k in ('pvf_bd', 'pvf_wt', 'vf_bd', 'vf_wt', 'ext_wit_bd', 'ext_wit_wt', 'wit_bd', 'wit_wt', 'wit_salt'):
        assert app[k] == oap[k]


def test_wire_format_oracle_roundtrip_and_layout():
    """oracle/wire.py: value i sits at bit offset i*bits, LSB first; out-of-range values are flagged."""
    import numpy as np
    import wire as owire
    x = np.zeros((2, 256), dtype=np.int16)
    x[0, 0], x[0, 1], x[0, 255] = -945, 945, 3
    p, ok = owire.pack(x, 11, 945)
    assert p.shape == (2, 352) and ok.tolist() == [1, 1]
    word = int.from_bytes(bytes(p[0, :4]), 'little')
    assert word & 0x7FF == 0 and (word >> 11) & 0x7FF == 1890
    tail = int.from_bytes(bytes(p[0, -4:]), 'little')
    assert tail >> (32 - 11) == 948
    assert np.array_equal(owire.unpack(p, 11, 945), x)
    x[1, 9] = 1103                                   # 1103 + 945 = 2^11
    assert owire.pack(x, 11, 945)[1].tolist() == [1, 0]
    rng = np.random.default_rng(0)
    u = rng.integers(0, 11777, size=(4, 2, 256)).astype(np.uint16)
    assert np.array_equal(owire.unpack(owire.pack(u, 14, 0)[0], 14, 0, np.uint16), u)
