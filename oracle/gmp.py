"""
GMP (``libgmp.so.10``) through ``ctypes``: the "gmpy2 path" of the reference.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``gmpy2.powmod`` / ``gmpy2.invert`` -- what the reference's third-party ``pow_mod`` / ``mod_inv``
call when the ``[gmpy]`` extra is installed (``pyproject.toml:44-49``) -- are thin wrappers over
``mpz_powm`` / ``mpz_invert``.  gmpy2 itself is not installable offline, so these are called
directly.  ``powm_batch_threads`` drives ``oracle/_build/libgmp_batch.so`` (C, pthreads; built by
``__graft_entry__.build()`` from ``oracle/c/gmp_batch.c``) for the timed CPU baseline.
"""

from __future__ import annotations

import ctypes
import os
from typing import Sequence

import numpy as np

_gmp = None


class _Mpz(ctypes.Structure):
    _fields_ = [
        ("_mp_alloc", ctypes.c_int),
        ("_mp_size", ctypes.c_int),
        ("_mp_d", ctypes.c_void_p),
    ]


def _lib():
    global _gmp
    if _gmp is None:
        g = ctypes.CDLL("libgmp.so.10")
        g.__gmpz_init.argtypes = [ctypes.POINTER(_Mpz)]
        g.__gmpz_clear.argtypes = [ctypes.POINTER(_Mpz)]
        g.__gmpz_set_str.argtypes = [ctypes.POINTER(_Mpz), ctypes.c_char_p, ctypes.c_int]
        g.__gmpz_get_str.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_Mpz)]
        g.__gmpz_get_str.restype = ctypes.c_char_p
        g.__gmpz_sizeinbase.argtypes = [ctypes.POINTER(_Mpz), ctypes.c_int]
        g.__gmpz_sizeinbase.restype = ctypes.c_size_t
        g.__gmpz_powm.argtypes = [ctypes.POINTER(_Mpz)] * 4
        g.__gmpz_invert.argtypes = [ctypes.POINTER(_Mpz)] * 3
        g.__gmpz_invert.restype = ctypes.c_int
        g.__gmpz_jacobi.argtypes = [ctypes.POINTER(_Mpz)] * 2
        g.__gmpz_jacobi.restype = ctypes.c_int
        # plain aliases: "__name" attribute access is name-mangled inside class bodies
        for name in ("init", "clear", "set_str", "get_str", "sizeinbase", "powm", "invert", "jacobi"):
            setattr(g, "z_" + name, getattr(g, "__gmpz_" + name))
        _gmp = g
    return _gmp


class Mpz:
    def __init__(self, value: int = 0) -> None:
        self.g = _lib()
        self.z = _Mpz()
        self.g.z_init(ctypes.byref(self.z))
        if value:
            self.set(value)

    def set(self, value: int) -> None:
        self.g.z_set_str(ctypes.byref(self.z), format(value, "x").encode(), 16)

    def get(self) -> int:
        size = self.g.z_sizeinbase(ctypes.byref(self.z), 16) + 2
        buf = ctypes.create_string_buffer(size)
        self.g.z_get_str(buf, 16, ctypes.byref(self.z))
        return int(buf.value, 16)

    def __del__(self) -> None:
        try:
            self.g.z_clear(ctypes.byref(self.z))
        except Exception:
            pass


def powm(base: int, exponent: int, modulus: int) -> int:
    """``mpz_powm`` with gmpy2.powmod's convention for negative exponents (invert first)."""
    g = _lib()
    if exponent < 0:
        base = invert(base, modulus)
        exponent = -exponent
    b, e, m, r = Mpz(base), Mpz(exponent), Mpz(modulus), Mpz()
    g.__gmpz_powm(ctypes.byref(r.z), ctypes.byref(b.z), ctypes.byref(e.z), ctypes.byref(m.z))
    return r.get()


def invert(value: int, modulus: int) -> int:
    g = _lib()
    v, m, r = Mpz(value % modulus), Mpz(modulus), Mpz()
    ok = g.__gmpz_invert(ctypes.byref(r.z), ctypes.byref(v.z), ctypes.byref(m.z))
    if not ok:
        raise ZeroDivisionError("invert() no inverse exists")
    return r.get()


def jacobi(a: int, n: int) -> int:
    g = _lib()
    za, zn = Mpz(a % n), Mpz(n)
    return int(g.__gmpz_jacobi(ctypes.byref(za.z), ctypes.byref(zn.z)))


# ---------------------------------------------------------------------------------------------
# threaded batch (C harness) -- the timed CPU baseline
# ---------------------------------------------------------------------------------------------

_BATCH_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libgmp_batch.so")
_batch = None


def batch_lib_path() -> str:
    return _BATCH_LIB


def _batch_lib():
    global _batch
    if _batch is None:
        if not os.path.exists(_BATCH_LIB):
            raise FileNotFoundError(
                f"{_BATCH_LIB} not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
            )
        lib = ctypes.CDLL(_BATCH_LIB)
        lib.gmp_powm_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.POINTER(ctypes.c_double),
        ]
        lib.gmp_powm_batch.restype = ctypes.c_int
        lib.gmp_powm_grouped.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_double),
        ]
        lib.gmp_powm_grouped.restype = ctypes.c_int
        _batch = lib
    return _batch


def powm_batch_threads(
    bases: np.ndarray, modulus_limbs: np.ndarray, exponent_limbs: np.ndarray, negative: bool,
    threads: int,
) -> tuple[np.ndarray, float]:
    """``out[i] = bases[i] ^ (+-exp) mod modulus`` for ``bases`` of shape [B, L] (uint32 LE limbs)
    with ``threads`` pthreads each looping ``mpz_powm`` (+ ``mpz_invert`` if negative).  Returns
    (out [B, L] uint32, seconds spent inside the threaded region)."""
    lib = _batch_lib()
    bases = np.ascontiguousarray(bases, dtype=np.uint32)
    B, L = bases.shape
    mod = np.ascontiguousarray(modulus_limbs, dtype=np.uint32)
    exp = np.ascontiguousarray(exponent_limbs, dtype=np.uint32)
    out = np.zeros((B, L), dtype=np.uint32)
    secs = ctypes.c_double(0.0)
    rc = lib.gmp_powm_batch(
        bases.ctypes.data, out.ctypes.data, B, L, mod.ctypes.data, mod.size,
        1 if negative else 0, exp.ctypes.data, exp.size, threads, ctypes.byref(secs),
    )
    if rc not in (0, 1):  # 1 = some base was not invertible (its row is all 0xff)
        raise RuntimeError(f"gmp_powm_batch failed rc={rc}")
    return out, secs.value


def powm_grouped_threads(
    bases: np.ndarray, moduli: np.ndarray, exps: np.ndarray, threads: int
) -> tuple[np.ndarray, float]:
    """Grouped variant: ``bases`` [G, K, L], ``moduli`` [G, L], ``exps`` [G, Le]."""
    lib = _batch_lib()
    bases = np.ascontiguousarray(bases, dtype=np.uint32)
    G, K, L = bases.shape
    moduli = np.ascontiguousarray(moduli, dtype=np.uint32)
    exps = np.ascontiguousarray(exps, dtype=np.uint32)
    out = np.zeros((G, K, L), dtype=np.uint32)
    secs = ctypes.c_double(0.0)
    rc = lib.gmp_powm_grouped(
        bases.ctypes.data, out.ctypes.data, G, K, L, moduli.ctypes.data, exps.ctypes.data,
        exps.shape[1], threads, ctypes.byref(secs),
    )
    if rc != 0:
        raise RuntimeError(f"gmp_powm_grouped failed rc={rc}")
    return out, secs.value


def int_to_limbs(value: int, limbs: int) -> np.ndarray:
    return np.frombuffer(value.to_bytes(4 * limbs, "little"), dtype=np.uint32).copy()


def ints_to_limbs(values: Sequence[int], limbs: int) -> np.ndarray:
    buf = b"".join(v.to_bytes(4 * limbs, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint32).reshape(len(values), limbs).copy()


def limbs_to_ints(arr: np.ndarray) -> list[int]:
    arr = np.ascontiguousarray(arr, dtype=np.uint32)
    width = arr.shape[-1] * 4
    raw = arr.tobytes()
    return [int.from_bytes(raw[i : i + width], "little") for i in range(0, len(raw), width)]
