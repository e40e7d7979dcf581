"""Worst-case magnitudes through the lazy-reduction kernels (run with -m gpu).

k_verify, k_sign, k_matvec and k_aggv_partial never reduce inside a transform: values grow by a few q per butterfly stage,
the FP32-assisted quotient is exact only below 2^22, products are accumulated in 64 bits and reduced once.  Random data sits
far from those limits, so the tests here drive every input the ABI admits to its extreme - coefficients all at +-q//2 or at
the int16 limits, NTT slots all q-1, the public row at +-q//2, the largest and smallest NTT-friendly moduli below 2^16, the
longest vectors - and compare with an independent numpy statement (schoolbook negacyclic products on int64).  The verify
cases are built so that the equation HOLDS (vk_right := key_ch * sig - vk_left * c): one wrong residue anywhere flips the
verdict, and a second pass with one tampered coefficient must flip it back to 0.
Reference semantics: lm_one_time_sigs.py:163-191 (sign, verify), :95-96 (key_ch * sk), bklm_one_time_agg_sigs.py:99-116."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

D = 256


def _ntt_primes():
    ps = [q for q in range(513, 65536, 512) if all(q % f for f in range(2, int(q ** 0.5) + 1))]
    return [ps[0], 11777, 39937, ps[-1]]            # smallest, the two shipped ones, the largest below 2^16


def negacyclic_mul(a, b, q):
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))
    res = full[:D].copy()
    res[:D - 1] -= full[D:]
    return res % q


def centred(x, q):
    x = np.asarray(x, dtype=np.int64) % q
    return np.where(x > (q - 1) // 2, x - q, x)


def dense_challenge(pairs):
    c = np.zeros(D, dtype=np.int64)
    for idx, s in pairs:
        c[int(idx)] = int(s)
    return c


def patterns(rng, shape, m):
    """Arrays of the given shape whose every entry has magnitude m."""
    alt = np.where(np.arange(shape[-1]) % 2 == 0, m, -m)
    yield 'all_plus', np.full(shape, m, dtype=np.int64)
    yield 'all_minus', np.full(shape, -m, dtype=np.int64)
    yield 'alternating', np.broadcast_to(alt, shape).astype(np.int64)
    yield 'random_signs', rng.choice(np.array([-m, m], dtype=np.int64), shape)


@pytest.mark.parametrize('q', _ntt_primes())
@pytest.mark.parametrize('l', [1, 13, 64])
def test_verify_holds_at_extreme_magnitudes(q, l):
    from lattice_cryptography_b200 import Engine, make_scheme
    rng = np.random.default_rng(1000 * l + q)
    half = q // 2
    sch = make_scheme(sk_bd=1, sk_wt=D, ch_bd=1, ch_wt=20)
    e = Engine(128, q, D, l)
    try:
        for kname, key_ch in patterns(rng, (l, D), half):
            e.set_key_ch(np.ascontiguousarray(key_ch.astype(np.int16)))
            for m in (half, 32767, 32768):
                sigs, names = [], []
                for sname, s in patterns(rng, (l, D), min(m, 32767)):
                    if m == 32768:
                        if sname != 'all_minus':
                            continue
                        s = np.full((l, D), -32768, dtype=np.int64)         # the one int16 value without a positive twin
                    sigs.append(s)
                    names.append(sname)
                n = len(sigs)
                sig = np.ascontiguousarray(np.stack(sigs).astype(np.int16))
                msgs = [bytes(rng.integers(0, 256, 40 + i, dtype=np.uint8)) for i in range(n)]
                pairs = e.challenge(sch, msgs)
                # left key halves: NTT slots all q-1 / coefficient form all +-q//2 / random
                vkl_ntt = np.empty((n, D), dtype=np.uint16)
                vkl_ntt[0] = q - 1
                if n > 1:
                    vkl_ntt[1:] = e.ntt_fwd(np.ascontiguousarray(rng.choice(np.array([-half, half]), (n - 1, D)).astype(np.int16)))
                vkl_coef = e.ntt_inv(vkl_ntt).astype(np.int64)
                vkr_coef = np.empty((n, D), dtype=np.int64)
                for i in range(n):
                    lhs = np.zeros(D, dtype=np.int64)
                    for j in range(l):
                        lhs = (lhs + negacyclic_mul(key_ch[j], sig[i, j], q)) % q
                    vkr_coef[i] = centred(lhs - negacyclic_mul(vkl_coef[i], dense_challenge(pairs[i, :, :]), q), q)
                vk = np.ascontiguousarray(np.stack([vkl_ntt, e.ntt_fwd(np.ascontiguousarray(vkr_coef.astype(np.int16)))], axis=1))
                bd = 32767 if m >= 32767 else half
                got = e.lm_verify(sch, vk, msgs, sig, bd, D)
                if m == 32768:
                    # |-32768| exceeds every bound the ABI can express (bd <= 32767): rejected by the bound test
                    assert got.tolist() == [0] * n, (kname, m)
                    continue
                assert got.tolist() == [1] * n, (kname, m, names, got.tolist())
                # one tampered coefficient of the right key half, and a bound one below the magnitude
                vkr_bad = vkr_coef.copy()
                vkr_bad[:, 255] = centred(vkr_bad[:, 255] + 1, q)
                vk_bad = np.ascontiguousarray(np.stack([vkl_ntt, e.ntt_fwd(np.ascontiguousarray(vkr_bad.astype(np.int16)))], axis=1))
                assert e.lm_verify(sch, vk_bad, msgs, sig, bd, D).tolist() == [0] * n, (kname, m)
                assert e.lm_verify(sch, vk, msgs, sig, min(m, 32767) - 1, D).tolist() == [0] * n, (kname, m)
    finally:
        e.close()


@pytest.mark.parametrize('q', _ntt_primes())
@pytest.mark.parametrize('l', [1, 23])
def test_sign_at_extreme_key_slots(q, l):
    """k_sign with every NTT slot of both key halves at q-1, at 0, and at random choices of the two."""
    from lattice_cryptography_b200 import Engine, make_scheme
    rng = np.random.default_rng(7 * l + q)
    e = Engine(128, q, D, l)
    try:
        e.set_key_ch(np.ascontiguousarray(rng.integers(-(q // 2), q // 2 + 1, (l, D)).astype(np.int16)))
        for ch_wt in (1, 20, 256):
            sch = make_scheme(sk_bd=1, sk_wt=D, ch_bd=1, ch_wt=ch_wt)
            sk = np.stack([np.full((2, l, D), q - 1), np.zeros((2, l, D)), rng.choice(np.array([0, q - 1]), (2, l, D)),
                           rng.integers(0, q, (2, l, D))]).astype(np.uint16)
            n = sk.shape[0]
            msgs = [bytes(rng.integers(0, 256, 60 + 3 * i, dtype=np.uint8)) for i in range(n)]
            pairs = e.challenge(sch, msgs)
            sig = e.lm_sign(sch, np.ascontiguousarray(sk), msgs).astype(np.int64)
            coef = e.ntt_inv(np.ascontiguousarray(sk.reshape(-1, D))).astype(np.int64).reshape(n, 2, l, D)
            for i in range(n):
                c = dense_challenge(pairs[i])
                for j in range(l):
                    want = centred(negacyclic_mul(coef[i, 0, j], c, q) + coef[i, 1, j], q)
                    assert np.array_equal(sig[i, j], want), (ch_wt, i, j)
    finally:
        e.close()


@pytest.mark.parametrize('q', _ntt_primes())
@pytest.mark.parametrize('l', [2, 64])
def test_keygen_product_at_extreme_public_row(q, l):
    """k_matvec: key_ch at +-q//2 times signing keys sampled with the widest bound a 16-bit context admits (sk_bd = q//2,
    every position non-zero): vk == key_ch * sk on int64."""
    from lattice_cryptography_b200 import Engine, make_scheme
    rng = np.random.default_rng(11 * l + q)
    half = q // 2
    sch = make_scheme(sk_bd=half, sk_wt=D, ch_bd=1, ch_wt=20)
    e = Engine(128, q, D, l)
    try:
        for kname, key_ch in patterns(rng, (l, D), half):
            e.set_key_ch(np.ascontiguousarray(key_ch.astype(np.int16)))
            seeds = [''.join(rng.choice(['0', '1'], 128)) for _ in range(3)]
            sk_coef, sk_ntt, vk_ntt, vk_coef = e.lm_keygen(sch, seeds)
            assert np.abs(sk_coef.astype(np.int64)).max() <= half and np.count_nonzero(sk_coef) == sk_coef.size
            for i in range(len(seeds)):
                for side in (0, 1):
                    want = np.zeros(D, dtype=np.int64)
                    for j in range(l):
                        want = (want + negacyclic_mul(key_ch[j], sk_coef[i, side, j], q)) % q
                    assert np.array_equal(vk_coef[i, side].astype(np.int64), centred(want, q)), (kname, i, side)
            assert np.array_equal(e.ntt_inv(np.ascontiguousarray(vk_ntt.reshape(-1, D))).reshape(vk_coef.shape), vk_coef)
            assert np.array_equal(e.ntt_fwd(np.ascontiguousarray(sk_coef.reshape(-1, D))).reshape(sk_ntt.shape), sk_ntt)
    finally:
        e.close()


@pytest.mark.parametrize('secpar,q,l', [(128, 11777, 13), (256, 39937, 23), (128, 64513, 2)])
def test_aggverify_partial_at_extreme_key_slots(secpar, q, l):
    """k_aggv_partial: sum_i (vk_left_i * c_i + vk_right_i) * ag_i with every key slot at q-1 (and 0, and mixed), rotations
    0 / 255 and both signs, counts around the persistent grid's trip boundaries; the int32 partial is congruent mod q to
    the numpy sum in NTT form."""
    from lattice_cryptography_b200 import Engine, make_scheme
    rng = np.random.default_rng(q + l)
    ch_wt = 20 if secpar == 128 else 50
    sch = make_scheme(sk_bd=1, sk_wt=D, ch_bd=1, ch_wt=ch_wt)
    e = Engine(secpar, q, D, l)
    try:
        e.set_key_ch(np.ascontiguousarray(rng.integers(-(q // 2), q // 2 + 1, (l, D)).astype(np.int16)))
        for count in (1, 2, 17, 600):
            vk = np.empty((count, 2, D), dtype=np.uint16)
            vk[:] = q - 1
            if count > 2:
                vk[1] = 0
                vk[2:] = rng.choice(np.array([0, q - 1], dtype=np.uint16), (count - 2, 2, D))
            msgs = [bytes(rng.integers(0, 256, 30 + (i % 50), dtype=np.uint8)) for i in range(count)]
            ks = rng.integers(0, D, count).astype(np.int16)
            ss = rng.choice(np.array([-1, 1], dtype=np.int16), count)
            ks[0] = 255
            if count > 1:
                ks[1] = 0
            ag = np.ascontiguousarray(np.stack([ks, ss], axis=1)[:, None, :])
            got = e.aggverify_partial(sch, vk, msgs, ag).astype(np.int64)
            pairs = e.challenge(sch, msgs)
            coef = e.ntt_inv(np.ascontiguousarray(vk.reshape(-1, D))).astype(np.int64).reshape(count, 2, D)
            total = np.zeros(D, dtype=np.int64)
            for i in range(count):
                t = (negacyclic_mul(coef[i, 0], dense_challenge(pairs[i]), q) + coef[i, 1]) % q
                mono = np.zeros(D, dtype=np.int64)
                mono[int(ks[i])] = int(ss[i])
                total = (total + negacyclic_mul(t, mono, q)) % q
            want = e.ntt_fwd(np.ascontiguousarray(centred(total, q).astype(np.int16)[None]))[0].astype(np.int64)
            assert np.array_equal((got - want) % q, np.zeros(D, dtype=np.int64)), count
    finally:
        e.close()
