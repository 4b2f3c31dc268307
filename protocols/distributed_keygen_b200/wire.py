"""
Wire format of the batched partial-decryption message, limb rows <-> bytes without going through
Python ints.

The reference broadcasts ``{"content": "partial_decryption_sequence", "value": [int, ...]}``
(``distributed_keygen.py:476-484``) and reads it back at ``:497-505``.  Its serializer (third-party
``tno.mpc.communication``, (or)msgpack based; not vendored) writes every integer that does not fit
64 bits as the map ``{"type": "int", "data": <little-endian two's-complement bytes>}`` with
``len(data) == (bit_length + 8) // 8`` -- the same tagging as in the stored-key blobs, which is
what pins it here (``tests/test_wire.py`` compares against the bytes of the reference's 24
fixtures).  The envelope the transport adds around the payload (message ids, compression) belongs
to the communication library and is out of scope; this module covers the payload body.

The codec itself is native (``csrc/dkg_wire.cu``: ``dkg_wire_encode_rows`` /
``dkg_wire_decode_rows``), one memcpy per element instead of one Python int per element.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._native import DkgError, check, lib

CONTENT = "partial_decryption_sequence"
_MSG_HEAD = b"\x82\xa7content" + bytes([0xA0 + len(CONTENT)]) + CONTENT.encode() + b"\xa5value"


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def encode_int_rows(rows: np.ndarray, prefix: bytes = b"") -> bytes:
    """``prefix`` + msgpack array of integers (big ones tagged as the reference does) from
    [count][limbs] rows."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    if rows.ndim != 2 or rows.shape[1] < 1:
        raise ValueError("rows must be [count][limbs]")
    count, limbs = rows.shape
    bound = 5 + count * (15 + 5 + 4 * limbs + 1)  # array32 header + tag + bin32 header + sign byte
    out = np.empty(len(prefix) + bound, dtype=np.uint8)
    out[: len(prefix)] = np.frombuffer(prefix, dtype=np.uint8)
    written = ctypes.c_size_t(0)
    check(lib.dkg_wire_encode_rows(_ptr(rows), count, limbs, ctypes.c_void_p(out.ctypes.data + len(prefix)),
                                   bound, ctypes.byref(written)))
    return out[: len(prefix) + written.value].tobytes()


def decode_int_rows(buf: bytes, limbs: int, offset: int = 0) -> tuple[np.ndarray, int]:
    """Inverse of :func:`encode_int_rows`: ([count][limbs] uint32, offset after the array).
    ``ValueError`` on a negative value, one that does not fit ``limbs`` limbs, or malformed input."""
    arr = np.frombuffer(buf, dtype=np.uint8)[offset:]
    if arr.size == 0:
        raise ValueError("empty buffer")
    count = ctypes.c_size_t(0)
    used = ctypes.c_size_t(0)
    try:
        check(lib.dkg_wire_decode_rows(_ptr(arr), arr.size, limbs, None, 0, ctypes.byref(count), ctypes.byref(used)))
        rows = np.empty((count.value, limbs), dtype=np.uint32)
        check(lib.dkg_wire_decode_rows(_ptr(arr), arr.size, limbs, _ptr(rows), count.value,
                                       ctypes.byref(count), ctypes.byref(used)))
    except (DkgError, MemoryError, OverflowError) as exc:
        raise ValueError(str(exc)) from None
    return rows, offset + used.value


def pack_partial_decryption_message(rows: np.ndarray) -> bytes:
    """Body of the broadcast at ``distributed_keygen.py:476-484`` from partial-decryption limb rows
    (what ``PaillierSharedKey.partial_decrypt_limbs`` returns)."""
    return encode_int_rows(rows, _MSG_HEAD)


def unpack_partial_decryption_message(buf: bytes, limbs: int) -> np.ndarray:
    """Limb rows from a received message body (``distributed_keygen.py:497-505``); the content
    field is checked like the reference's assertion at ``:500-502``."""
    buf = bytes(buf)
    if buf[: len(_MSG_HEAD)] != _MSG_HEAD:
        import msgpack

        msg = msgpack.unpackb(buf, raw=False, strict_map_key=False)
        content = msg.get("content") if isinstance(msg, dict) else None
        raise AssertionError(f"received a share for {content}, but expected {CONTENT}")
    rows, end = decode_int_rows(buf, limbs, len(_MSG_HEAD))
    if end != len(buf):
        raise ValueError("trailing bytes after the partial decryption list")
    return rows
