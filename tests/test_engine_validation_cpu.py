"""Host-side argument checks of the array-level API (no GPU needed): the C ABI trusts the extents it is given,
so lattice_cryptography_b200.engine must refuse short, mistyped or mis-shaped buffers before they reach it."""
import numpy as np
import pytest

from lattice_cryptography_b200.engine import Engine, _lead, _want


def test_want_accepts_exact_shape_and_dtype():
    x = np.zeros((3, 13, 256), dtype=np.int16)
    assert _want(x, 'sig', np.int16, (3, 13, 256)) is x
    assert _want(x, 'sig', (np.int16, np.uint16), (None, 13, 256)) is x
    assert _lead(x, 'sig', np.int16, 256) == 39
    assert _lead(np.zeros(256, dtype=np.int16), 'p', np.int16, 256) == 1


@pytest.mark.parametrize('bad', [
    np.zeros((3, 12, 256), dtype=np.int16),          # a short signature: one polynomial missing
    np.zeros((3, 13, 255), dtype=np.int16),
    np.zeros((2, 13, 256), dtype=np.int16),          # row count differs from the message count
    np.zeros((3, 13, 256), dtype=np.int32),          # wrong element width
    np.zeros((3, 13 * 256), dtype=np.int16),
    [[0] * 256] * 13,                                # not an array at all
    None,
])
def test_want_rejects(bad):
    with pytest.raises(ValueError):
        _want(bad, 'sig', np.int16, (3, 13, 256))


def test_ragged_tuple_is_checked():
    blob = np.zeros(10, dtype=np.uint8)
    Engine._rag((blob, np.array([0, 4, 10], dtype=np.int64)))
    for off in (np.array([0, 4, 11], dtype=np.int64), np.array([1, 4, 10], dtype=np.int64),
                np.array([0, 6, 4, 10], dtype=np.int64), np.array([0, 4, 10], dtype=np.int32), np.zeros(0, dtype=np.int64)):
        with pytest.raises(ValueError):
            Engine._rag((blob, off))
    with pytest.raises(ValueError):
        Engine._rag((blob.astype(np.int8), np.array([0, 10], dtype=np.int64)))


def test_torch_tensors_are_checked_too():
    torch = pytest.importorskip('torch')
    t = torch.zeros((2, 2, 256), dtype=torch.uint16)
    assert _want(t, 'vk', np.uint16, (2, 2, 256)) is t
    with pytest.raises(ValueError):
        _want(t.view(torch.int16), 'vk', np.uint16, (2, 2, 256))
    with pytest.raises(ValueError):
        _want(t[:1], 'vk', np.uint16, (2, 2, 256))
