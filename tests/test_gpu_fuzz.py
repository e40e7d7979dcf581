"""Randomised differential test of the LM-OTS path (run with -m gpu): random NTT-friendly moduli below 2^16, vector
lengths, security parameters, key / challenge bounds and weights, batch sizes, seed and message lengths - keygen, sign and
verify (with tampered signatures) through the C ABI against oracle/lcb_oracle.c, bit for bit.

The default budget is a few seconds with a fixed seed (part of the suite); LCB_FUZZ_SECONDS / LCB_FUZZ_SEED extend it:
    LCB_FUZZ_SECONDS=240 LCB_FUZZ_SEED=$RANDOM python -m pytest tests/test_gpu_fuzz.py -m gpu -q -s
Reference semantics: lm_one_time_sigs.py:64-97 (keys), :141-170 (challenge, sign), :173-191 (verify)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

D = 256


def _primes():
    return [q for q in range(513, 65536, 512) if all(q % f for f in range(2, int(q ** 0.5) + 1))]


def test_random_parameter_sets_vs_c_oracle():
    import c_oracle
    from lattice_cryptography_b200 import Engine, make_scheme, ragged
    budget = float(os.environ.get('LCB_FUZZ_SECONDS', '12'))
    seed = int(os.environ.get('LCB_FUZZ_SEED', '20261018'))
    rng = np.random.default_rng(seed)
    primes = _primes()
    t_end = time.monotonic() + budget
    rounds = keys = sigs = verdicts = 0
    while time.monotonic() < t_end or rounds < 3:
        q = int(rng.choice(primes))
        l = int(rng.integers(1, 25))
        secpar = int(rng.choice([64, 128, 192, 256]))
        sk_bd = int(rng.integers(1, min(200, q // 2) + 1))
        ch_wt = int(rng.integers(1, 61))
        ch_bd = int(rng.choice([1, 1, 1, 2, 5]))
        n = int(rng.integers(1, 160))
        vf_bd = min(q // 2, sk_bd * (1 + ch_wt * ch_bd))
        case = dict(seed=seed, round=rounds, q=q, l=l, secpar=secpar, sk_bd=sk_bd, ch_wt=ch_wt, ch_bd=ch_bd, n=n)
        key_ch = rng.integers(-(q // 2), q // 2 + 1, size=(l, D)).astype(np.int16)
        e = Engine(secpar, q, D, l)
        try:
            e.set_key_ch(key_ch)
            sch = make_scheme(sk_bd=sk_bd, sk_wt=D, ch_bd=ch_bd, ch_wt=ch_wt)
            p = c_oracle.params(secpar, q, l, sk_bd, ch_wt, ch_bd=ch_bd)
            seeds = [''.join(rng.choice(['0', '1'], secpar + int(rng.integers(0, 150)))) for _ in range(n)]
            chmsgs = [bytes(rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8)) for _ in range(n)]
            sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
            sig = e.lm_sign(sch, sk_ntt, chmsgs)
            for i in sorted(set(int(x) for x in rng.integers(0, n, 3))):
                skl, skr, vkl, vkr = c_oracle.lm_keygen(p, key_ch, seeds[i].encode())
                assert np.array_equal(skl, sk_coef[i, 0]) and np.array_equal(skr, sk_coef[i, 1]), case
                assert np.array_equal(vkl, vk_coef[i, 0]) and np.array_equal(vkr, vk_coef[i, 1]), case
                assert np.array_equal(c_oracle.lm_sign(p, skl, skr, chmsgs[i]), sig[i]), case
                keys += 1
                sigs += 1
            bad = sig.copy()
            for i in range(0, n, 3):
                kind = int(rng.integers(0, 3))
                j, k = int(rng.integers(0, l)), int(rng.integers(0, D))
                if kind == 0:
                    bad[i, j, k] += int(rng.choice([-1, 1]))
                elif kind == 1:
                    bad[i, j, k] = vf_bd + 1 if vf_bd < 32767 else -32768           # over the bound
                else:                                    # another representative of the same residue, where int16 has one
                    x = int(bad[i, j, k])
                    alt = x + q if x + q <= 32767 else (x - q if x - q >= -32768 else x + 1)
                    bad[i, j, k] = alt
            blob, off = ragged(chmsgs)
            got = e.lm_verify(sch, vk_ntt, (blob, off), bad, vf_bd, D)
            want = c_oracle.lm_verify_batch(p, key_ch, vk_coef, blob, off, bad, vf_bd, D)
            assert np.array_equal(got, want), (case, np.flatnonzero(got != want)[:8].tolist())
            honest = e.lm_verify(sch, vk_ntt, (blob, off), sig, vf_bd, D)
            assert np.array_equal(honest, c_oracle.lm_verify_batch(p, key_ch, vk_coef, blob, off, sig, vf_bd, D)), case
            verdicts += 2 * n
        finally:
            e.close()
        rounds += 1
    print(f'fuzz: seed {seed}, {rounds} parameter sets, {keys} keys, {sigs} signatures, {verdicts} verdicts compared')
