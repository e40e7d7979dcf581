"""CPU ORACLE (test infrastructure): generates tests/golden/*.npz + *.json by running the
REFERENCE'S OWN scheme modules (imported unmodified from /root/reference through
oracle/ref_loader.py) on top of the restated lattice_algebra.

Run in the build container only:   python oracle/gen_golden.py
The fixtures travel to the GPU box; /root/reference does not.

Recorded per case: the exact hash-input strings the reference built (they contain CPython
object addresses), the decoded (index, coefficient) draw sequences, dense
centred coefficient arrays of every key / challenge / signature, and the verdicts
(including tampered inputs).  Seeds follow tests/test_lm_one_time_sigs.py:64
(`bin(j)[2:].zfill(secpar)`), messages follow benchmarks/demo_signing.py:5,
tests/test_adaptor_sigs.py:204 and benchmarks/benchmark_lm_one_time_sigs.py:75.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from ref_loader import load_reference  # noqa: E402

otk_mod, lm, bklm, ad = load_reference()      # real lattice_algebra underneath if reachable (oracle/l1.py)
import lattice_algebra as la  # noqa: E402  (the restatement; ref_loader put it on sys.path)
import schemes  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
KEY_CH_SEED = 'lcb200 golden key_ch v1'
SEED_INTS = [0, 1, 2, 12345]
MESSAGES = ['QRL is awesome!', 'Blessed are the cheesemakers.']


def dense(p):
    return np.array(schemes.dense_of_poly(p), dtype=np.int16)


def dense_vec(v):
    return np.array(schemes.dense_of_vec(v), dtype=np.int16)


def draw_pairs(secpar, lp, salt, msg, bd, wt, length=None):
    """(index, coef) pairs in draw order, straight from the L1 decoder."""
    bti, btd = la.bits_to_indices(secpar, lp.degree, wt), la.bits_to_decode(secpar, bd)
    nb = la.get_gen_bytes_per_poly(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {'bd': bd, 'wt': wt}, wt, bti, btd)
    n = 1 if length is None else length
    bits = la.binary_digest(msg, nb * n, salt)
    out = []
    for i in range(n):
        cd = la.decode2polycoefs(secpar, lp, la.UNIFORM_INFINITY_WEIGHT, {'bd': bd, 'wt': wt},
                                 bits[i * 8 * nb:(i + 1) * 8 * nb], wt, bti, btd)
        out.append([[k, v] for k, v in cd.items()])
    return np.array(out if length is not None else out[0], dtype=np.int16)


def fix_key_ch(pp, secpar):
    key_ch = schemes.key_ch_from_seed(secpar, KEY_CH_SEED)
    pp['scheme_parameters'].key_ch = key_ch
    return key_ch


def main():
    # "l1": which lattice_algebra the reference's modules ran on when these vectors were made.  Anything but
    # "lattice_algebra==0.1.1" means PARITY UNPINNED: the vectors then pin the engine to the
    # restatement only.
    arrays, meta = {}, {'key_ch_seed': KEY_CH_SEED, 'l1': ref_loader.L1_LABEL, 'cases': {}}
    for secpar in (128, 256):
        tag = f's{secpar}'
        # ---------------------------------------------------------------- LM-OTS
        pp = lm.make_setup_parameters(secpar)
        key_ch = fix_key_ch(pp, secpar)
        lp = pp['scheme_parameters'].lp
        arrays[f'{tag}_key_ch'] = dense_vec(key_ch)
        arrays[f'{tag}_key_ch_pairs'] = draw_pairs(secpar, lp, 'KEY_CH_SEED', KEY_CH_SEED, lp.modulus // 2,
                                                  lp.degree, lp.length)
        m = {'q': lp.modulus, 'd': lp.degree, 'l': lp.length, 'vf_bd': pp['vf_bd'], 'vf_wt': pp['vf_wt'],
             'sk_bd': pp['sk_bd'], 'sk_wt': pp['sk_wt'], 'ch_wt': pp['ch_wt'], 'rou': lp.rou, 'lm': [], }
        seeds = [bin(j)[2:].zfill(secpar) for j in SEED_INTS]
        seed_objs = [otk_mod.SecretSeed(secpar=secpar, lp=lp, seed=s) for s in seeds]
        keys = lm.keygen_core(pp=pp, num_keys_to_gen=len(seeds), seeds=seed_objs)
        msgs = MESSAGES + [bin(0xC0FFEE + j)[2:].zfill(secpar) for j in range(len(seeds) - len(MESSAGES))]
        for j, (key, msg) in enumerate(zip(keys, msgs)):
            _, sk, vk = key
            sig = lm.sign(pp=pp, otk=key, msg=msg)
            chmsg = str(vk) + ', ' + msg
            c = lm.make_signature_challenge(pp=pp, otvk=vk, msg=msg)
            ok = lm.verify(pp=pp, otvk=vk, msg=msg, sig=sig)
            assert ok
            bad_msg_ok = lm.verify(pp=pp, otvk=vk, msg=msg + '!', sig=sig)
            pre = f'{tag}_lm{j}'
            arrays[f'{pre}_skL'], arrays[f'{pre}_skR'] = dense_vec(sk[0]), dense_vec(sk[1])
            arrays[f'{pre}_vkL'], arrays[f'{pre}_vkR'] = dense(vk[0]), dense(vk[1])
            arrays[f'{pre}_c'] = dense(c)
            arrays[f'{pre}_c_pairs'] = draw_pairs(secpar, lp, pp['ch_salt'], chmsg, pp['ch_bd'], pp['ch_wt'])
            arrays[f'{pre}_sig'] = dense_vec(sig)
            if j == 0:
                arrays[f'{pre}_skL_pairs'] = draw_pairs(secpar, lp, pp['sk_salt'] + 'LEFT', seeds[j], pp['sk_bd'],
                                                        pp['sk_wt'], lp.length)
            # tampered signature: one coefficient +1 (fails the equation), one coefficient out of bound
            t1 = arrays[f'{pre}_sig'].copy()
            t1[j % lp.length, (7 * j + 3) % lp.degree] += 1
            t2 = arrays[f'{pre}_sig'].copy()
            t2[(j + 1) % lp.length, (11 * j + 5) % lp.degree] = pp['vf_bd'] + 1
            v1 = lm.verify(pp=pp, otvk=vk, msg=msg, sig=schemes.vec_from_dense(lp, t1.tolist()))
            v2 = lm.verify(pp=pp, otvk=vk, msg=msg, sig=schemes.vec_from_dense(lp, t2.tolist()))
            arrays[f'{pre}_sig_t1'], arrays[f'{pre}_sig_t2'] = t1, t2
            digest = hashlib.shake_256((pp['sk_salt'] + 'LEFT' + seeds[j]).encode()).digest(4096)
            m['lm'].append({'seed': seeds[j], 'msg': msg, 'chmsg': chmsg, 'verdict': ok,
                            'verdict_bad_msg': bad_msg_ok, 'verdict_t1': v1, 'verdict_t2': v2,
                            'skL_digest_first64': digest[:64].hex(),
                            'skL_digest4096_sha256': hashlib.sha256(digest).hexdigest()})
        # ---------------------------------------------------------------- BKLM
        ppb = bklm.make_setup_parameters(secpar)
        ppb['scheme_parameters'].key_ch = key_ch
        m['bklm'] = []
        for cap in (2, 4):
            ppb['ag_cap'] = cap
            ppb['avf_wt'] = max(1, min(lp.degree, cap * ppb['ag_wt'] * ppb['vf_wt']))
            ppb['avf_bd'] = max(1, min(lp.modulus // 2, cap * min(ppb['ag_wt'], ppb['vf_wt']) * ppb['ag_bd'] *
                                       ppb['vf_bd']))
            ks = keys[:cap]
            bmsgs = [bin(0xABCDEF01 + 17 * j)[2:].zfill(32)[-32:] for j in range(cap)]
            sigs = [lm.sign(pp=ppb, otk=k, msg=mm) for k, mm in zip(ks, bmsgs)]
            vks = [k[2] for k in ks]
            ag_sig = bklm.aggregate(pp=ppb, otvks=vks, msgs=bmsgs, sigs=sigs)
            ok = bklm.aggregate_verify(pp=ppb, otvks=vks, msgs=bmsgs, ag_sig=ag_sig)
            assert ok
            srt_keys, srt_msgs = bklm.prepare_make_agg_coefs(otvks=vks, msgs=bmsgs)
            order = [vks.index(k) for k in srt_keys]
            agmsg = str(list(zip(srt_keys, srt_msgs)))
            coefs = bklm.make_agg_coefs(pp=ppb, otvks=vks, msgs=bmsgs)
            pre = f'{tag}_bk{cap}'
            arrays[f'{pre}_sigs'] = np.stack([dense_vec(s) for s in sigs])
            arrays[f'{pre}_ag_coefs'] = np.stack([dense(a) for a in coefs])
            arrays[f'{pre}_ag_sig'] = dense_vec(ag_sig)
            bad = arrays[f'{pre}_ag_sig'].copy()
            bad[0, 0] += 1
            bad_ok = bklm.aggregate_verify(pp=ppb, otvks=vks, msgs=bmsgs,
                                           ag_sig=schemes.vec_from_dense(lp, bad.tolist()))
            m['bklm'].append({'cap': cap, 'avf_bd': ppb['avf_bd'], 'avf_wt': ppb['avf_wt'], 'msgs': bmsgs,
                              'key_index': list(range(cap)), 'sorted_order': order, 'agmsg': agmsg,
                              'chmsgs': [str(k) + ', ' + mm for k, mm in zip(vks, bmsgs)],
                              'verdict': ok, 'verdict_tampered': bad_ok})
        # ---------------------------------------------------------------- adaptor
        ppa = ad.make_setup_parameters(secpar)
        ppa['scheme_parameters'].key_ch = key_ch
        m['adaptor_params'] = {k: ppa[k] for k in ('pvf_bd', 'pvf_wt', 'vf_bd', 'vf_wt', 'ext_wit_bd',
                                                   'ext_wit_wt', 'wit_bd', 'wit_wt')}
        m['adaptor'] = []
        akeys = ad.keygen(pp=ppa, num_keys_to_gen=2, seeds=seed_objs[:2])
        wit_seeds = [bin(777 + j)[2:].zfill(secpar) for j in range(2)]
        wits = ad.witgen(pp=ppa, num_wits_to_gen=2,
                         seeds=[otk_mod.SecretSeed(secpar=secpar, lp=lp, seed=s) for s in wit_seeds])
        for j, (key, ws) in enumerate(zip(akeys, wits)):
            _, wit, st = ws
            msg = MESSAGES[1]
            chmsg = str(st) + ', ' + str(key[2]) + ', ' + msg
            presig = ad.presign(pp=ppa, otk=key, msg=msg, st=st)
            pv = ad.preverify(pp=ppa, otvk=key[2], msg=msg, st=st, presig=presig)
            sig = ad.adapt(presig=presig, wit=wit)
            vv = ad.verify(pp=ppa, otvk=key[2], msg=msg, st=st, sig=sig)
            ext = ad.extract(pp=ppa, presig=presig, sig=sig)
            wv = ad.witness_verify(pp=ppa, wit=ext, st=st)
            pv_on_sig = ad.preverify(pp=ppa, otvk=key[2], msg=msg, st=st, presig=sig)  # adapted sig is no presig
            assert pv and vv and wv
            pre = f'{tag}_ad{j}'
            arrays[f'{pre}_wit'], arrays[f'{pre}_st'] = dense_vec(wit.key), dense(st.key)
            arrays[f'{pre}_wit_pairs'] = draw_pairs(secpar, lp, ppa['wit_salt'], wit_seeds[j], ppa['wit_bd'],
                                                    ppa['wit_wt'], lp.length)
            arrays[f'{pre}_presig'], arrays[f'{pre}_sig'] = dense_vec(presig), dense_vec(sig)
            arrays[f'{pre}_ext'] = dense_vec(ext.key)
            m['adaptor'].append({'key_index': j, 'wit_seed': wit_seeds[j], 'msg': msg, 'chmsg': chmsg,
                                 'preverify': pv, 'verify': vv, 'witness_verify': wv,
                                 'preverify_of_adapted': pv_on_sig})
        meta['cases'][str(secpar)] = m
    np.savez_compressed(os.path.join(OUT, 'golden.npz'), **arrays)
    with open(os.path.join(OUT, 'golden.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('wrote', len(arrays), 'arrays;', os.path.getsize(os.path.join(OUT, 'golden.npz')), 'bytes npz')


if __name__ == '__main__':
    main()
