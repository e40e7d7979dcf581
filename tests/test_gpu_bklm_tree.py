"""BKLM, linear-time NON-REFERENCE mode (pp['ag_mode'] = 'tree', SURVEY.md section 8(f)4 last clause): the aggregation
message is bound into a two-level SHAKE256 tree commitment and every coefficient hashes ag_salt || str(i) || commitment.
Checked against hashlib (commitment and coefficients), against a numpy statement of the aggregate, and for the
properties a mode switch must have (round trip, rejection across modes, sharding)."""
import hashlib
from secrets import randbits

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tree_commitment(raw: bytes) -> str:
    """hashlib restatement of bklm_one_time_agg_sigs.commit_aggregation_message."""
    leaves = b''.join(hashlib.shake_256(raw[a:a + 8192]).digest(32) for a in range(0, len(raw), 8192))
    return hashlib.shake_256(b'LCB200-AGTREE' + len(raw).to_bytes(8, 'little') + leaves).digest(32).hex()


@pytest.mark.parametrize('nbytes', [0, 1, 135, 136, 8191, 8192, 8193, 16384, 100_000, 1_000_003])
def test_commitment_matches_hashlib(nbytes):
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    pp = bk.make_setup_parameters(128)
    raw = np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8).tobytes()
    assert bk.commit_aggregation_message(pp, raw) == tree_commitment(raw)


def test_mode_switch_is_validated():
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    pp = bk.make_setup_parameters(128)
    assert pp.get('ag_mode', 'reference') == 'reference'
    with pytest.raises(ValueError, match='ag_mode'):
        bk.set_aggregation_mode(pp, 'fast')


@pytest.mark.parametrize('secpar', [128, 256])
def test_tree_mode_round_trip_and_coefficients(secpar):
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import keygen, sign
    n = 40
    pp = bk.set_aggregation_mode(bk.set_aggregation_capacity(bk.make_setup_parameters(secpar), n), 'tree')
    ref_pp = bk.set_aggregation_capacity(bk.make_setup_parameters(secpar), n)
    ref_pp['scheme_parameters'] = pp['scheme_parameters']           # same key_ch: only the mode differs
    keys = keygen(pp=pp, num_keys_to_gen=n)
    msgs = [bin(randbits(32))[2:].zfill(32) for _ in keys]
    sigs = [sign(pp=pp, otk=k, msg=m) for k, m in zip(keys, msgs)]
    vks = [k[2] for k in keys]
    srt_keys, srt_msgs, srt_sigs = bk.prepare_aggregate(vks, msgs, sigs)
    agmsg = str(list(zip(srt_keys, srt_msgs)))
    root = tree_commitment(agmsg.encode())
    coefs = bk.make_agg_coefs(pp=pp, otvks=vks, msgs=msgs)
    want = []
    for i in range(n):
        dg = hashlib.shake_256(b'AG_SALT' + str(i).encode() + root.encode()).digest(2)
        want.append((dg[0], 1 if dg[1] & 0x80 else -1))
        cr, norm, wt = coefs[i].get_coef_rep()
        assert cr == {want[i][0]: want[i][1]} and norm == 1 and wt == 1
    # aggregate == sum of signed negacyclic rotations of the sorted signatures, centred
    lp = pp['scheme_parameters'].lp
    q = lp.modulus
    acc = np.zeros((lp.length, 256), dtype=np.int64)
    for s, (k, sg) in zip(srt_sigs, want):
        c = s.coef.astype(np.int64)
        acc += sg * np.concatenate([-c[:, 256 - k:], c[:, :256 - k]], axis=1)
    acc = (acc % q + q) % q
    acc = np.where(acc > q // 2, acc - q, acc)
    ag_sig = bk.aggregate(pp=pp, otvks=vks, msgs=msgs, sigs=sigs)
    assert np.array_equal(ag_sig.coef, acc)
    assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs, ag_sig=ag_sig) is True
    assert bk.aggregate_verify(pp=pp, otvks=vks[::-1], msgs=msgs[::-1], ag_sig=ag_sig) is True
    assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs[1:] + msgs[:1], ag_sig=ag_sig) is False
    # the two modes are different schemes: neither accepts the other's aggregate
    assert bk.aggregate_verify(pp=ref_pp, otvks=vks, msgs=msgs, ag_sig=ag_sig) is False
    ref_sig = bk.aggregate(pp=ref_pp, otvks=vks, msgs=msgs, sigs=sigs)
    assert ref_sig != ag_sig
    assert bk.aggregate_verify(pp=ref_pp, otvks=vks, msgs=msgs, ag_sig=ref_sig) is True
    assert bk.aggregate_verify(pp=pp, otvks=vks, msgs=msgs, ag_sig=ref_sig) is False


def test_tree_mode_sharded_equals_whole():
    from lattice_cryptography_b200 import bklm_one_time_agg_sigs as bk
    from lattice_cryptography_b200.lm_one_time_sigs import challenge_messages, keygen, sign
    n = 30
    pp = bk.set_aggregation_mode(bk.set_aggregation_capacity(bk.make_setup_parameters(128), n), 'tree')
    keys = keygen(pp=pp, num_keys_to_gen=n)
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
    for a, b in ((0, 11), (11, 12), (12, 30)):
        part += bk.aggregate_shard(pp, np.ascontiguousarray(sig_arr[a:b]), agmsg, a)
        vpart += bk.aggregate_verify_shard(pp, np.ascontiguousarray(vk_arr[a:b]), chm[a:b], agmsg, a)
    ag = bk.aggregate_finish(pp, part.astype(np.int32))
    assert np.array_equal(ag, whole.coef)
    assert bk.aggregate_verify_finish(pp, vpart.astype(np.int32), ag, n)
