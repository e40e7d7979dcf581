"""CPU tests pinning the plain-C oracle (oracle/lcb_oracle.c: schoolbook products, bit-by-bit decoder)
to hashlib, to the Python restatement the reference's own modules run on, and to tests/golden/."""
import hashlib

import numpy as np
import pytest

import c_oracle
import lattice_algebra as la
import schemes

SHIPPED = {128: dict(q=11777, l=13, sk_bd=45, ch_wt=20), 256: dict(q=39937, l=23, sk_bd=65, ch_wt=50)}


def test_shake256_against_hashlib():
    rng = np.random.default_rng(3)
    for n in [0, 1, 135, 136, 137, 272, 500]:
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        for out in (1, 136, 137, 1000):
            assert c_oracle.shake256(data, out) == hashlib.shake_256(data).digest(out)


@pytest.mark.parametrize('secpar,bd,wt,vec_len', [(128, 45, 256, 2), (256, 1, 50, 1), (128, 5888, 256, 1), (128, 3, 7, 3)])
def test_decoder_against_python_restatement(secpar, bd, wt, vec_len):
    lp = schemes.lattice_parameters(SHIPPED[secpar]['q'], 256, vec_len)
    bti, btd = la.bits_to_indices(secpar, 256, wt), la.bits_to_decode(secpar, bd)
    nb = la.get_gen_bytes_per_poly(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {}, wt, bti, btd)
    for msg in ['', 'abc', '0110' * 40]:
        dense, pairs = c_oracle.hash2polyvec(secpar, 256, 'SOME_SALT', msg.encode(), bd, wt, vec_len)
        bits = la.binary_digest(msg, nb * vec_len, 'SOME_SALT')
        for v in range(vec_len):
            cd = la.decode2polycoefs(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {'bd': bd, 'wt': wt},
                                     bits[v * 8 * nb:(v + 1) * 8 * nb], wt, bti, btd)
            assert [[int(a), int(b)] for a, b in pairs[v]] == [[k, c] for k, c in cd.items()]


@pytest.mark.parametrize('secpar', [128, 256])
def test_lm_against_golden(golden, secpar):
    arrays, meta = golden
    m = meta['cases'][str(secpar)]
    s = SHIPPED[secpar]
    p = c_oracle.params(secpar, s['q'], s['l'], s['sk_bd'], s['ch_wt'])
    key_ch = np.ascontiguousarray(arrays[f's{secpar}_key_ch'])
    for j, case in enumerate(m['lm']):
        pre = f's{secpar}_lm{j}'
        skl, skr, vkl, vkr = c_oracle.lm_keygen(p, key_ch, case['seed'].encode())
        assert np.array_equal(skl, arrays[pre + '_skL']) and np.array_equal(skr, arrays[pre + '_skR'])
        assert np.array_equal(vkl, arrays[pre + '_vkL']) and np.array_equal(vkr, arrays[pre + '_vkR'])
        sig = c_oracle.lm_sign(p, skl, skr, case['chmsg'].encode())
        assert np.array_equal(sig, arrays[pre + '_sig'])
        assert c_oracle.lm_verify(p, key_ch, vkl, vkr, case['chmsg'].encode(), sig, m['vf_bd'], m['vf_wt']) == case['verdict']
        assert c_oracle.lm_verify(p, key_ch, vkl, vkr, (case['chmsg'] + '!').encode(), sig, m['vf_bd'], m['vf_wt']) == \
            case['verdict_bad_msg']
        for tag in ('t1', 't2'):
            assert c_oracle.lm_verify(p, key_ch, vkl, vkr, case['chmsg'].encode(), arrays[f'{pre}_sig_{tag}'],
                                      m['vf_bd'], m['vf_wt']) == case[f'verdict_{tag}']
    for case in m['bklm']:
        want = arrays[f's{secpar}_bk{case["cap"]}_ag_coefs']
        for i in range(case['cap']):
            k, sgn = c_oracle.agg_coef(secpar, 256, 'AG_SALT', i, case['agmsg'].encode())
            assert want[i, k] == sgn and np.count_nonzero(want[i]) == 1
    ad = m['adaptor'][0]
    wit, _ = c_oracle.hash2polyvec(secpar, 256, 'WIT_SALT', ad['wit_seed'].encode(), 1, 20, s['l'])
    assert np.array_equal(wit, arrays[f's{secpar}_ad0_wit'])
