"""Host-only behaviour of the drop-in objects (no GPU needed: nothing here multiplies or transforms):
deepcopy / pickle round trips (the reference deep-copies polynomials in every operator and pickles key
tuples across its multiprocessing Pool, lm_one_time_sigs.py:100-123), equality, container indexing and the
batched form of the unseeded keygen path (make_random_seed, lm_one_time_sigs.py:58-61)."""
import copy
import pickle

import numpy as np


def _lp():
    from lattice_cryptography_b200.lattice_algebra import LatticeParameters
    return LatticeParameters(modulus=11777, degree=256, length=13)


def test_polynomial_objects_deepcopy_and_pickle():
    from lattice_cryptography_b200.lattice_algebra import Polynomial, PolynomialVector
    lp = _lp()
    f = Polynomial(lp=lp, coefs={0: 1, 7: -5888, 255: 5888}, const_time_flag=False)
    g = Polynomial(lp=lp, coefs={3: 2})
    v = PolynomialVector(lp=lp, entries=[f] + [g] * 12)
    for obj in (lp, f, v):
        for clone in (copy.deepcopy(obj), pickle.loads(pickle.dumps(obj))):
            assert clone == obj and clone is not obj
    c = copy.deepcopy(f)
    assert c.get_coef_rep() == ({0: 1, 7: -5888, 255: 5888}, 5888, 3) and c.const_time_flag is False
    c._coef[0] = 2                                   # a deep copy owns its storage
    assert f.get_coef_rep()[0][0] == 1 and c != f


def test_key_objects_pickle_and_index():
    from lattice_cryptography_b200.lattice_algebra import Polynomial, PolynomialVector
    from lattice_cryptography_b200.one_time_keys import OneTimeSigningKey, OneTimeVerificationKey, SecretSeed
    lp = _lp()
    seed = SecretSeed(secpar=128, lp=lp, seed='01' * 64)
    half = PolynomialVector(lp=lp, entries=[Polynomial(lp=lp, coefs={i: i + 1}) for i in range(13)])
    sk = OneTimeSigningKey(secpar=128, lp=lp, left_key=half, right_key=copy.deepcopy(half))
    vk = OneTimeVerificationKey(secpar=128, lp=lp, left_key=Polynomial(lp=lp, coefs={1: 1}),
                                right_key=Polynomial(lp=lp, coefs={2: -1}))
    seed2, sk2, vk2 = pickle.loads(pickle.dumps((seed, sk, vk)))
    assert seed2 == seed and sk2 == sk and vk2 == vk
    assert sk2[0] == sk.left_key and vk2[1] == vk.right_key
    # as in the reference, str() of a key is its identity, so the copy is a different signer ...
    assert str(vk2) != str(vk) and ' object at 0x' in str(vk)
    # ... unless the opt-in content-based form is switched on
    from lattice_cryptography_b200 import one_time_keys as otk
    otk.set_content_str(True)
    try:
        assert str(vk2) == str(vk) == repr(vk) and str(vk).startswith('<OneTimeVerificationKey ')
        other = OneTimeVerificationKey(secpar=128, lp=lp, left_key=Polynomial(lp=lp, coefs={1: 1}),
                                       right_key=Polynomial(lp=lp, coefs={2: 1}))
        assert str(other) != str(vk) and str([vk])[1:-1] == str(vk)
    finally:
        otk.set_content_str(False)
    assert ' object at 0x' in str(vk)


def test_wire_widths_and_ragged_helpers():
    """Host-only arithmetic of the wire format (bits per coefficient / slot) and of the ragged-blob helper."""
    from lattice_cryptography_b200 import ragged, wire

    class _LP:
        def __init__(self, q):
            self.modulus = q

    class _SP:
        def __init__(self, q):
            self.lp = _LP(q)
    assert wire.sig_bits({'vf_bd': 945}) == 11 and wire.sig_bits({'vf_bd': 3315}) == 13
    assert wire.sig_bits({'vf_bd': 1023}) == 11 and wire.sig_bits({'vf_bd': 1024}) == 12      # 2*bd + 1 values
    assert wire.sig_bits({'pvf_bd': 945, 'vf_bd': 946}, bound_key='pvf_bd') == 11
    assert wire.key_bits({'scheme_parameters': _SP(11777)}) == 14 and wire.key_bits({'scheme_parameters': _SP(39937)}) == 16
    blob, off = ragged(['ab', b'', 'cé', b'\x00\x01\x02'])
    assert off.tolist() == [0, 2, 2, 5, 8] and bytes(blob) == b'abc\xc3\xa9\x00\x01\x02'
    blob, off = ragged([])
    assert off.tolist() == [0] and blob.size >= 0
