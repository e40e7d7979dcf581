"""Device-side expansion of host entropy for the unseeded keygen path (second half;
make_random_seed, lm_one_time_sigs.py:58-61): one 32-byte secret -> SHAKE256(secret || le64(i)) on the GPU -> N seed
bitstrings.  Checked against hashlib, for reproducibility of key i from seed i, and for shape / alphabet / balance."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _want(secret, i, secpar):
    dg = hashlib.shake_256(secret + i.to_bytes(8, 'little')).digest(secpar // 8)
    return ''.join(format(b, '08b') for b in dg)


@pytest.mark.parametrize('secpar', [128, 256])
def test_expansion_matches_hashlib(secpar):
    from lattice_cryptography_b200 import Engine
    q, l = {128: (11777, 13), 256: (39937, 23)}[secpar]
    e = Engine(secpar, q, 256, l)
    secret = bytes(range(7, 39))
    blob, off = e.expand_seeds(secret, 300, first=2 ** 33 - 5)
    assert blob.dtype == np.uint8 and blob.shape == (300 * secpar,) and set(np.unique(blob)) <= {48, 49}
    assert off.tolist() == [i * secpar for i in range(301)]
    rows = blob.reshape(300, secpar)
    for i in (0, 1, 4, 5, 6, 150, 299):                  # the counter crosses 2^33 inside the batch
        assert bytes(rows[i]).decode() == _want(secret, 2 ** 33 - 5 + i, secpar), i
    # device-resident output, same bytes
    d_blob, d_off = e.expand_seeds(secret, 300, first=2 ** 33 - 5, device=True)
    assert np.array_equal(d_blob.cpu().numpy(), blob) and d_off.cpu().numpy().tolist() == off.tolist()
    with pytest.raises(ValueError):
        e.expand_seeds(b'short', 4)
    e.close()


def test_random_seed_batch_shape_balance_and_reproducibility():
    from lattice_cryptography_b200 import lm_one_time_sigs as lm
    pp = lm.make_setup_parameters(128)
    blob, off = lm.random_seed_batch(pp, 4000)
    assert off.dtype == np.int64 and off[0] == 0 and off[-1] == 512000 and (np.diff(off) == 128).all()
    rows = blob.reshape(4000, 128)
    assert len({bytes(r) for r in rows}) == 4000 and 0.49 < (rows == 49).mean() < 0.51
    assert (np.abs((rows == 49).mean(axis=0) - 0.5) < 0.05).all()           # every bit position is balanced
    other, _ = lm.random_seed_batch(pp, 16)
    assert not np.array_equal(other.reshape(16, 128), rows[:16])           # a fresh secret per call
    # a caller-held secret makes the batch reproducible, and seed i reproduces key i through the seeded path
    a, _ = lm.random_seed_batch(pp, 6, secret=b'\x01' * 32)
    b, _ = lm.random_seed_batch(pp, 6, secret=b'\x01' * 32)
    assert np.array_equal(a, b)
    fresh = lm.keygen_batch(pp, (a, np.arange(7, dtype=np.int64) * 128))
    one = lm.keygen_batch(pp, [bytes(a.reshape(6, 128)[3]).decode()])
    assert np.array_equal(fresh['vk_ntt'][3], one['vk_ntt'][0]) and np.array_equal(fresh['sk_coef'][3], one['sk_coef'][0])
