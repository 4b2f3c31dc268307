"""Python int <-> little-endian uint32 limb arrays (the C ABI's number format, which is also the
byte order of ``int.to_bytes(..., "little")`` and of the reference's stored keys)."""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np


def limbs_for_bits(bits: int) -> int:
    return max(1, (bits + 31) // 32)


def int_to_limbs(value: int, limbs: int) -> np.ndarray:
    return np.frombuffer(value.to_bytes(4 * limbs, "little"), dtype=np.uint32).copy()


def ints_to_limbs(values: Sequence[int] | Iterable[int], limbs: int) -> np.ndarray:
    values = list(values)
    width = 4 * limbs
    buf = bytearray(width * len(values))
    for i, v in enumerate(values):
        buf[i * width : (i + 1) * width] = v.to_bytes(width, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint32).reshape(len(values), limbs).copy()


def limbs_to_int(arr: np.ndarray) -> int:
    return int.from_bytes(np.ascontiguousarray(arr, dtype=np.uint32).tobytes(), "little")


def limbs_to_ints(arr: np.ndarray) -> list[int]:
    arr = np.ascontiguousarray(arr, dtype=np.uint32)
    width = arr.shape[-1] * 4
    raw = arr.tobytes()
    return [int.from_bytes(raw[i : i + width], "little") for i in range(0, len(raw), width)]
