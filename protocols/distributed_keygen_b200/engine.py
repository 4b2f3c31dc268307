"""
Thin object layer over the C ABI: one ``ModexpContext`` per (modulus, signed exponent) -- i.e. per
key -- and batched calls on numpy limb arrays or Python ints.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import _native
from .limbs import int_to_limbs, ints_to_limbs, limbs_for_bits, limbs_to_ints


class ModexpContext:
    """``out[i] = bases[i] ** exponent mod modulus`` for a fixed odd modulus and a fixed signed
    exponent: the batched form of ``pow_mod`` (+ ``mod_inv`` for negative exponents) at
    ``paillier_shared_key.py:89-92`` of the reference."""

    def __init__(self, modulus: int, exponent: int, device: int = 0) -> None:
        if modulus <= 0 or modulus % 2 == 0:
            raise ValueError("modulus must be a positive odd integer")
        self.modulus = modulus
        self.exponent = exponent
        self.device = device
        self.limbs = limbs_for_bits(modulus.bit_length())
        mag = abs(exponent)
        exp_limbs = limbs_for_bits(mag.bit_length())
        self._mod = int_to_limbs(modulus, self.limbs)
        self._exp = int_to_limbs(mag, exp_limbs)
        handle = ctypes.c_void_p()
        _native.check(
            _native.lib.dkg_modexp_ctx_create(
                device, self._mod.ctypes.data, self.limbs, self._exp.ctypes.data, exp_limbs,
                1 if exponent < 0 else 0, ctypes.byref(handle),
            )
        )
        self._h = handle

    def close(self) -> None:
        if getattr(self, "_h", None):
            _native.lib.dkg_modexp_ctx_destroy(self._h)
            self._h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict[str, int]:
        arr = (ctypes.c_int * 8)()
        _native.check(_native.lib.dkg_modexp_ctx_info(self._h, ctypes.byref(arr)))
        keys = ["K", "M", "padded_limbs", "window_bits", "windows", "exponent_bits", "warps_per_cta", "ctas"]
        return dict(zip(keys, list(arr)))

    def modexp_limbs(self, bases: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """bases: uint32 [count, limbs] (values < modulus).  Returns (out [count, limbs], status)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint32)
        if bases.ndim != 2 or bases.shape[1] != self.limbs:
            raise ValueError(f"bases must have shape [count, {self.limbs}]")
        count = bases.shape[0]
        out = np.zeros_like(bases)
        status = np.zeros(count, dtype=np.uint8)
        _native.check(
            _native.lib.dkg_modexp_batch(self._h, bases.ctypes.data, out.ctypes.data, status.ctypes.data, count)
        )
        return out, status

    def modexp_device(self, d_bases: int, d_out: int, d_status: int, count: int, stream: int) -> None:
        """Device-pointer variant (pointers and a cudaStream_t as integers); asynchronous."""
        _native.check(_native.lib.dkg_modexp_batch_device(self._h, d_bases, d_out, d_status, count, stream))

    def modexp(self, bases: Sequence[int]) -> list[int]:
        """Python-int convenience wrapper with the reference's error behaviour: a base that is not
        invertible under a negative exponent raises ``ZeroDivisionError`` (as ``mod_inv`` does)."""
        arr = ints_to_limbs([b % self.modulus for b in bases], self.limbs)
        out, status = self.modexp_limbs(arr)
        if status.any():
            raise ZeroDivisionError("base is not invertible for the given modulus")
        return limbs_to_ints(out)


def launch_count() -> int:
    return int(_native.lib.dkg_launch_count())
