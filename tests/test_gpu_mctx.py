"""The multi-device context of the C ABI (lcb_mctx_*, SURVEY.md 8(b) `devices[], ndev`): one call shards a host batch
over several contexts and reduces the BKLM partial sums on the devices.  On a 1-GPU box the device list repeats
ordinal 0 (two or three contexts, one host thread each, on the same GPU), on a multi-GPU box it uses every GPU;
either way the results must equal the single-context ones bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices(k):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(k)]


@pytest.mark.parametrize('parts', [1, 2, 3])
@pytest.mark.parametrize('secpar,q,d,l,sk_bd,ch_wt', [(128, 11777, 256, 13, 45, 20), (128, 12289, 512, 2, 3, 6)])
def test_sharded_calls_equal_single_context(parts, secpar, q, d, l, sk_bd, ch_wt):
    from lattice_cryptography_b200 import Engine, MultiEngine, make_scheme
    n = 50                                                    # not a multiple of 3: ragged shards
    one = Engine(secpar, q, d, l)
    many = MultiEngine(secpar, q, d, l, _devices(parts))
    sch = make_scheme(sk_bd=sk_bd, sk_wt=d, ch_bd=1, ch_wt=ch_wt)
    key_ch, _ = one.hash2polyvec('KEY_CH_SEED', ['mctx'], q // 2, d, l)
    key_ch = np.ascontiguousarray(key_ch[0])
    one.set_key_ch(key_ch)
    many.set_key_ch(key_ch)
    seeds = [bin(7 + 13 * j)[2:].zfill(secpar - j % 3) + '0' * (j % 3) for j in range(n)]
    msgs = [bin(3 * j)[2:].zfill(32) for j in range(n)]
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * j:010x}>' for j in range(n)]
    chm = [k + ', ' + m + 'x' * (j % 4) for j, (k, m) in enumerate(zip(ident, msgs))]          # ragged hash inputs
    agmsg = ('[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']').encode()
    a = one.lm_keygen(sch, seeds)
    b = many.lm_keygen(sch, seeds)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    sig = one.lm_sign(sch, a[1], chm)
    assert np.array_equal(many.lm_sign(sch, b[1], chm), sig)
    vf_bd = min(q // 2, sk_bd * (1 + ch_wt))
    bad = sig.copy()
    bad[n - 1, 0, 0] += 1
    bad[17, l - 1, 5] = vf_bd + 1
    want = one.lm_verify(sch, a[2], chm, bad, vf_bd, d)
    assert np.array_equal(many.lm_verify(sch, b[2], chm, bad, vf_bd, d), want) and want.sum() == n - 2
    if d == 256:
        import math
        sb, kb = math.ceil(math.log2(2 * vf_bd + 1)), math.ceil(math.log2(q))
        sp, vp = one.pack(bad, sb, vf_bd), one.pack(a[2], kb, 0)
        assert np.array_equal(many.lm_verify_packed(sch, vp, kb, chm, sp, sb, vf_bd, vf_bd, d), want)
    # BKLM: whole aggregate / aggregate_verify in one call each
    coefs = one.agg_coefs(sch, agmsg, 0, n)
    ag_one = one.aggregate_finish(one.aggregate_partial(sch, sig, coefs))
    ag_many = many.bklm_aggregate(sch, sig, agmsg)
    assert np.array_equal(ag_many, ag_one)
    avf_bd = min(q // 2, n * vf_bd)
    assert many.bklm_aggregate_verify(sch, b[2], chm, agmsg, ag_many, n, avf_bd, d) is True
    tam = ag_many.copy()
    tam[0, 0] += 1
    assert many.bklm_aggregate_verify(sch, b[2], chm, agmsg, tam, n, avf_bd, d) is False
    assert many.bklm_aggregate_verify(sch, b[2], chm, agmsg, ag_many, n - 1, avf_bd, d) is False      # over capacity
    assert many.launch_count > 0
    # device pointers are refused: this entry shards host memory
    import torch
    from lattice_cryptography_b200 import LcbError
    with pytest.raises(LcbError):
        many.lm_sign(sch, torch.from_numpy(b[1].view(np.int16)).cuda().view(torch.uint16), chm)
    one.close()
    many.close()
