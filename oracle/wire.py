"""TEST INFRASTRUCTURE ONLY - numpy restatement of the packed wire format declared in include/lcb200.h
(lcb_pack_batch / lcb_unpack_batch).  The reference has no serialisation (one_time_keys.py:197-237), so
this file IS the specification the CUDA kernels in csrc/wire.cu are checked against; nothing in the
product imports it."""
import numpy as np


def pack(values: np.ndarray, bits: int, bias: int):
    """values int16/uint16 [..., 256] -> (uint8 [..., 32*bits], in_range uint8 [...])."""
    v = (values.astype(np.int64).astype(np.uint16).astype(np.int64) + bias) & 0xFFFF
    ok = (v < (1 << bits)).all(axis=-1).astype(np.uint8)
    v &= (1 << bits) - 1
    b = ((v[..., None] >> np.arange(bits)) & 1).astype(np.uint8)          # LSB first
    flat = b.reshape(values.shape[:-1] + (256 * bits,))
    return np.packbits(flat, axis=-1, bitorder='little'), ok


def unpack(packed: np.ndarray, bits: int, bias: int, dtype=np.int16):
    b = np.unpackbits(packed, axis=-1, bitorder='little').reshape(packed.shape[:-1] + (256, bits)).astype(np.int64)
    v = (b << np.arange(bits)).sum(axis=-1)
    return ((v - bias) & 0xFFFF).astype(np.uint16).view(np.uint16).astype(np.uint16).view(dtype) if dtype == np.int16 \
        else ((v - bias) & 0xFFFF).astype(np.uint16)
