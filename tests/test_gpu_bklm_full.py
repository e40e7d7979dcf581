"""BKLM aggregate / aggregate-verify at the BASELINE size (configs[3]: 2^16 signatures per aggregate, secpar 128),
checked against the oracle instead of against itself (a wrong-but-consistent coefficient derivation
cancels on both sides of key_ch*ag_sig == sum(...), so "verdict: true" proves nothing).

One process plays all 8 ranks (8 contiguous shards of the sorted list, partial sums added on the host exactly as
the NCCL reduce adds them), so the test runs on a 1-GPU box; tests/test_gpu_multi.py is the 2-rank NCCL variant.

  (a) aggregation coefficients: samples from every decimal-digit class, both ends of every shard and random
      positions, against hashlib.shake_256(b'AG_SALT' + str(i) + agmsg) on the 8.1 MB message
      (bklm_one_time_agg_sigs.py:60-81);
  (b) aggregate: the summed partials == a numpy statement of sum_i sig_i ** ag_i (bklm_one_time_agg_sigs.py:96)
      over all 2^16 signatures, fed with the coefficients checked in (a);
  (c) aggregate_verify: per shard, the partial sum over a 32-signature subsample == the C oracle's
      sum_i (vk_left_i * c_i + vk_right_i) * ag_i (schoolbook products, oracle challenges, oracle keys from the same
      seeds; bklm_one_time_agg_sigs.py:107-115); then the full verdict, and a tampered aggregate.
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOG2N, SHARDS, SECPAR = 16, 8, 128
Q, L, D, SK_BD, CH_WT, VF_BD = 11777, 13, 256, 45, 20, 945


def _agg_numpy(sigs, ks, ss):
    """sum_i s_i * X^k_i * sig_i in Z[X]/(X^256+1), all signatures at once (chunks bound the index arrays)."""
    acc = np.zeros((L, D), dtype=np.int64)
    p = np.arange(D)
    for a in range(0, len(sigs), 2048):
        k = ks[a:a + 2048, None].astype(np.int64)
        src = (p[None, :] - k) & (D - 1)                                    # out[p] = +-in[(p - k) mod d]
        sign = np.where(p[None, :] < k, -1, 1) * ss[a:a + 2048, None]
        rot = np.take_along_axis(sigs[a:a + 2048].astype(np.int64), np.broadcast_to(src[:, None, :], (src.shape[0], L, D)), axis=2)
        acc += (rot * sign[:, None, :]).sum(axis=0)
    return acc


@pytest.fixture(scope='module')
def world():
    import c_oracle
    from lattice_cryptography_b200 import Engine, make_scheme, ragged
    from lattice_cryptography_b200.distributed import shard_range
    total = 1 << LOG2N
    eng = Engine(SECPAR, Q, D, L)
    sch = make_scheme(sk_bd=SK_BD, sk_wt=D, ch_bd=1, ch_wt=CH_WT)
    key_ch, _ = eng.hash2polyvec('KEY_CH_SEED', ['bklm full-size parity'], Q // 2, D, L)
    key_ch = np.ascontiguousarray(key_ch[0])
    eng.set_key_ch(key_ch)
    rng = np.random.default_rng(20261018)
    bits = rng.integers(0, 2, (total, 32), dtype=np.uint8) + ord('0')
    msgs = [bytes(r).decode() for r in bits]
    # addresses ascend, so the reference's sort by str(otvk) is the identity
    ident = [f'<lattice_cryptography.one_time_keys.OneTimeVerificationKey object at 0x7f{16 * i:010x}>' for i in range(total)]
    agmsg = ('[' + ', '.join(f"({k}, '{m}')" for k, m in zip(ident, msgs)) + ']').encode()
    seeds = [bin((0x9E3779B97F4A7C15 * (i + 1)) % (1 << SECPAR))[2:].zfill(SECPAR) for i in range(total)]
    chm = [k + ', ' + m for k, m in zip(ident, msgs)]
    _, sk_ntt, vk_ntt, vk_coef = eng.lm_keygen(sch, seeds, want_sk_coef=False)
    sigs = eng.lm_sign(sch, sk_ntt, chm)
    del sk_ntt
    shards = [shard_range(total, r, SHARDS) for r in range(SHARDS)]
    coefs = np.concatenate([eng.agg_coefs(sch, agmsg, s, c) for s, c in shards])       # int16[N, 1, 2]
    yield dict(eng=eng, sch=sch, key_ch=key_ch, agmsg=agmsg, seeds=seeds, chm=chm, vk_ntt=vk_ntt, vk_coef=vk_coef,
               sigs=sigs, shards=shards, coefs=coefs, total=total, rng=rng, orc=c_oracle, ragged=ragged)
    eng.close()


def test_aggregation_coefficients_against_hashlib(world):
    total, coefs, agmsg = world['total'], world['coefs'], world['agmsg']
    assert len(agmsg) == 124 * total          # 122 per entry + ', ' separators + brackets
    picks = {0, 5, 9, 10, 11, 57, 99, 100, 101, 500, 999, 1000, 1001, 5000, 9999, 10000, 10001, 40000, total - 1}
    for s, c in world['shards']:
        picks |= {s, s + 1, s + c - 1}
    picks |= set(int(i) for i in world['rng'].integers(0, total, 40))
    assert len(picks) >= 64
    h0 = hashlib.shake_256(b'AG_SALT')
    for i in sorted(picks):
        h = h0.copy()
        h.update(str(i).encode() + agmsg)
        dg = h.digest(2)
        assert (int(coefs[i, 0, 0]), int(coefs[i, 0, 1])) == (dg[0], 1 if dg[1] & 0x80 else -1), f'coefficient {i}'
    # every digit-length class of the salt and every shard is represented
    assert {len(str(i)) for i in picks} == {1, 2, 3, 4, 5}


def test_aggregate_matches_numpy_sum_over_all_signatures(world):
    eng, sch, sigs, coefs = world['eng'], world['sch'], world['sigs'], world['coefs']
    part = np.zeros((L, D), dtype=np.int64)
    for s, c in world['shards']:
        p = eng.aggregate_partial(sch, np.ascontiguousarray(sigs[s:s + c]), np.ascontiguousarray(coefs[s:s + c]))
        assert p.dtype == np.int32 and int(np.abs(p.astype(np.int64)).max()) < 2 ** 31 // SHARDS     # 8 partials cannot overflow
        part += p
    ag = eng.aggregate_finish(part.astype(np.int32))
    want = _agg_numpy(sigs, coefs[:, 0, 0], coefs[:, 0, 1].astype(np.int64)) % Q
    want = np.where(want > (Q - 1) // 2, want - Q, want).astype(np.int16)
    assert np.array_equal(ag, want)
    world['ag_sig'] = ag


def test_aggregate_verify_partials_against_c_oracle_subsample(world):
    eng, sch, orc, coefs = world['eng'], world['sch'], world['orc'], world['coefs']
    p = orc.params(SECPAR, Q, L, SK_BD, CH_WT)
    for s, c in world['shards']:
        pick = np.sort(np.concatenate([[s, s + c - 1], s + world['rng'].choice(np.arange(1, c - 1), 30, replace=False)]))
        # oracle side: keys from the seeds, challenges from the hash inputs, schoolbook products
        acc = np.zeros(D, dtype=np.int64)
        for i in pick:
            _, _, vkl, vkr = orc.lm_keygen(p, world['key_ch'], world['seeds'][i].encode())
            assert np.array_equal(vkl, world['vk_coef'][i, 0]) and np.array_equal(vkr, world['vk_coef'][i, 1])
            ch, _ = orc.hash2polyvec(SECPAR, D, 'CH_SALT', world['chm'][i].encode(), 1, CH_WT, 1)
            t = orc.poly_mul(Q, vkl, ch[0]).astype(np.int64) + vkr
            t = ((t + Q // 2) % Q - Q // 2).astype(np.int16)
            mono = np.zeros(D, dtype=np.int16)
            mono[int(coefs[i, 0, 0])] = int(coefs[i, 0, 1])
            acc += orc.poly_mul(Q, t, mono)
        want = ((acc + Q // 2) % Q - Q // 2).astype(np.int16)
        # engine side: the same 32 items as one shard
        part = eng.aggverify_partial(sch, np.ascontiguousarray(world['vk_ntt'][pick]), [world['chm'][i] for i in pick],
                                     np.ascontiguousarray(coefs[pick]))
        got = eng.ntt_inv((part.astype(np.int64) % Q).astype(np.uint16)[None])[0]
        assert np.array_equal(got, want), f'shard starting at {s}'


def test_aggregate_verify_verdict_over_eight_shards(world):
    eng, sch, coefs, total = world['eng'], world['sch'], world['coefs'], world['total']
    ag = world.get('ag_sig')
    if ag is None:
        pytest.skip('needs the aggregate from test_aggregate_matches_numpy_sum_over_all_signatures')
    vpart = np.zeros(D, dtype=np.int64)
    for s, c in world['shards']:
        vpart += eng.aggverify_partial(sch, np.ascontiguousarray(world['vk_ntt'][s:s + c]), world['chm'][s:s + c],
                                       np.ascontiguousarray(coefs[s:s + c]))
    avf_bd = min(Q // 2, total * VF_BD)          # bklm_one_time_agg_sigs.py:36-43 with ag_cap = 2^16: clamps to q//2
    assert avf_bd == 5888
    assert eng.aggverify_finish(vpart.astype(np.int32), ag, total, total, avf_bd, D) is True
    bad = ag.copy()
    bad[3, 77] += 1
    assert eng.aggverify_finish(vpart.astype(np.int32), bad, total, total, avf_bd, D) is False
    assert eng.aggverify_finish(vpart.astype(np.int32), ag, total, total - 1, avf_bd, D) is False      # over capacity
    wrong = vpart.copy()
    wrong[5] += 1
    assert eng.aggverify_finish(wrong.astype(np.int32), ag, total, total, avf_bd, D) is False


@pytest.mark.parametrize('secpar,q,l', [(128, 11777, 13), (256, 39937, 23)])
def test_aggregate_partial_shapes_and_extremes(secpar, q, l):
    """k_agg_partial on synthetic rows: every count around its batching (4 rows per warp, 8 warps per block), rotations 0
    and 255, both signs, coefficients at the int16 extremes the verifier admits (|x| <= q // 2) - against numpy.  The kernel
    reads wrapped positions from a complemented copy of the row and repairs the sum with a per-rotation histogram: the
    cases here are the ones that repair has to get right (all rows wrapped / none wrapped / one row only)."""
    from lattice_cryptography_b200 import Engine, make_scheme
    eng = Engine(secpar, q, D, l)
    sch = make_scheme(sk_bd=1, sk_wt=D, ch_bd=1, ch_wt=20)
    rng = np.random.default_rng(secpar)
    for count in (1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 1000, 4099):
        sigs = rng.integers(-(q // 2), q // 2 + 1, (count, l, D), dtype=np.int16)
        ks = rng.integers(0, D, count).astype(np.int16)
        ss = rng.choice(np.array([-1, 1], dtype=np.int16), count)
        if count >= 3:
            ks[0], ks[1], ks[2] = 0, 255, 1
            sigs[1] = q // 2
            sigs[2] = -(q // 2)
        for variant in ('mixed', 'all_k255', 'all_k0', 'all_negative'):
            k2, s2 = ks.copy(), ss.copy()
            if variant == 'all_k255':
                k2[:] = 255
            elif variant == 'all_k0':
                k2[:] = 0
            elif variant == 'all_negative':
                s2[:] = -1
            pairs = np.ascontiguousarray(np.stack([k2, s2], axis=1)[:, None, :])
            got = eng.aggregate_partial(sch, sigs, pairs).astype(np.int64)
            p = np.arange(D)
            src = (p[None, :] - k2[:, None].astype(np.int64)) & (D - 1)
            sign = np.where(p[None, :] < k2[:, None], -1, 1) * s2[:, None].astype(np.int64)
            rot = np.take_along_axis(sigs.astype(np.int64), np.broadcast_to(src[:, None, :], (count, l, D)), axis=2)
            want = (rot * sign[:, None, :]).sum(axis=0)
            assert np.array_equal((got - want) % q, np.zeros_like(want)), (count, variant)
            fin = eng.aggregate_finish(got.astype(np.int32)).astype(np.int64)
            cen = want % q
            cen = np.where(cen > q // 2, cen - q, cen)
            assert np.array_equal(fin, cen), (count, variant)
    eng.close()
