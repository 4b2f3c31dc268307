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

    def __init__(self, modulus: int, exponent: int, device: int = 0, root: int | None = None) -> None:
        """``root``: if given, ``modulus`` must equal ``root ** 2``; the exponentiation then runs in
        pair arithmetic modulo ``root`` (same results, ~1.6x fewer multiplications)."""
        if modulus <= 0 or modulus % 2 == 0:
            raise ValueError("modulus must be a positive odd integer")
        if root is not None and root * root != modulus:
            raise ValueError("root ** 2 must equal the modulus")
        self.modulus = modulus
        self.exponent = exponent
        self.device = device
        self.limbs = limbs_for_bits(modulus.bit_length())
        mag = abs(exponent)
        exp_limbs = limbs_for_bits(mag.bit_length())
        self._mod = int_to_limbs(modulus, self.limbs)
        self._exp = int_to_limbs(mag, exp_limbs)
        handle = ctypes.c_void_p()
        if root is None:
            _native.check(
                _native.lib.dkg_modexp_ctx_create(
                    device, self._mod.ctypes.data, self.limbs, self._exp.ctypes.data, exp_limbs,
                    1 if exponent < 0 else 0, ctypes.byref(handle),
                )
            )
        else:
            root_limbs = limbs_for_bits(root.bit_length())
            self._root = int_to_limbs(root, root_limbs)
            _native.check(
                _native.lib.dkg_modexp_ctx_create_nsq(
                    device, self._root.ctypes.data, root_limbs, self._exp.ctypes.data, exp_limbs,
                    1 if exponent < 0 else 0, ctypes.byref(handle),
                )
            )
        self.root = root
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
        arr = (ctypes.c_int * 12)()
        _native.check(_native.lib.dkg_modexp_ctx_info(self._h, ctypes.byref(arr)))
        keys = ["K", "M", "padded_limbs", "window_bits", "windows", "exponent_bits", "warps_per_cta", "ctas",
                "pair_arithmetic", "pair_K", "pair_M", "pair_warps_per_cta"]
        out = dict(zip(keys, list(arr)))
        if out["pair_arithmetic"]:
            out["warps_per_cta"] = out["pair_warps_per_cta"]
        return out

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


class CombineContext:
    """Batched share combination: the arithmetic of ``PaillierSharedKey.decrypt``
    (``paillier_shared_key.py:108-125`` of the reference) for ``shares`` = degree + 1 partial
    decryptions per ciphertext."""

    def __init__(self, n: int, theta_inv: int, shares: int, device: int = 0) -> None:
        if n <= 0 or n % 2 == 0:
            raise ValueError("n must be a positive odd integer")
        self.n = n
        self.shares = shares
        self.device = device
        self.n_limbs = limbs_for_bits(n.bit_length())
        self.n2_limbs = limbs_for_bits((n * n).bit_length())
        self._n = int_to_limbs(n, self.n_limbs)
        self._th = int_to_limbs(theta_inv % n, self.n_limbs)
        handle = ctypes.c_void_p()
        _native.check(
            _native.lib.dkg_combine_ctx_create(
                device, self._n.ctypes.data, self.n_limbs, self._th.ctypes.data, shares, ctypes.byref(handle)
            )
        )
        self._h = handle
        assert _native.lib.dkg_combine_n2_limbs(self._h) == self.n2_limbs

    def close(self) -> None:
        if getattr(self, "_h", None):
            _native.lib.dkg_combine_ctx_destroy(self._h)
            self._h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def combine_limbs(self, partials: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """partials: uint32 [shares, count, n2_limbs].  Returns (plaintexts [count, n_limbs],
        status [count]) with status 2 where (x - 1) is not divisible by N."""
        partials = np.ascontiguousarray(partials, dtype=np.uint32)
        if partials.ndim != 3 or partials.shape[0] != self.shares or partials.shape[2] != self.n2_limbs:
            raise ValueError(f"partials must have shape [{self.shares}, count, {self.n2_limbs}]")
        count = partials.shape[1]
        out = np.zeros((count, self.n_limbs), dtype=np.uint32)
        status = np.zeros(count, dtype=np.uint8)
        _native.check(
            _native.lib.dkg_combine_batch(self._h, partials.ctypes.data, out.ctypes.data, status.ctypes.data, count)
        )
        return out, status

    def combine_device(self, d_partials: int, d_out: int, d_status: int, count: int, stream: int) -> None:
        _native.check(_native.lib.dkg_combine_batch_device(self._h, d_partials, d_out, d_status, count, stream))


class EncryptContext:
    """Batched Paillier encryption with g = N + 1 (``distributed_keygen.py:712``):
    ``(1 + m N) * r^N mod N^2`` -- the third-party ``Paillier.encrypt`` / ``randomize`` arithmetic --
    or just the randomness ``r^N mod N^2`` when no plaintexts are given."""

    def __init__(self, n: int, device: int = 0) -> None:
        self.n = n
        self.n_limbs = limbs_for_bits(n.bit_length())
        self._n = int_to_limbs(n, self.n_limbs)
        self._ctx = ModexpContext(n * n, n, device, root=n)
        self.n2_limbs = self._ctx.limbs

    def close(self) -> None:
        self._ctx.close()

    def encrypt_limbs(self, r: np.ndarray, m: np.ndarray | None) -> np.ndarray:
        r = np.ascontiguousarray(r, dtype=np.uint32)
        count = r.shape[0]
        if r.ndim != 2 or r.shape[1] != self.n_limbs:
            raise ValueError(f"r must have shape [count, {self.n_limbs}]")
        m_ptr = None
        if m is not None:
            m = np.ascontiguousarray(m, dtype=np.uint32)
            if m.shape != r.shape:
                raise ValueError("m must have the same shape as r")
            m_ptr = m.ctypes.data
        out = np.zeros((count, self.n2_limbs), dtype=np.uint32)
        _native.check(
            _native.lib.dkg_encrypt_batch(
                self._ctx._h, self._n.ctypes.data, self.n_limbs, r.ctypes.data, m_ptr, out.ctypes.data, count
            )
        )
        return out

    def encrypt(self, plaintexts: Sequence[int], randomness: Sequence[int]) -> list[int]:
        """Raw ciphertext values for already-encoded integer plaintexts (taken modulo N, so negative
        values are ``N - |m|`` as in the reference's encoding)."""
        m = ints_to_limbs([p % self.n for p in plaintexts], self.n_limbs)
        r = ints_to_limbs([x % self.n for x in randomness], self.n_limbs)
        return limbs_to_ints(self.encrypt_limbs(r, m))

    def randomness(self, randomness: Sequence[int]) -> list[int]:
        """``r^N mod N^2`` (what the reference's randomness pre-generation computes)."""
        r = ints_to_limbs([x % self.n for x in randomness], self.n_limbs)
        return limbs_to_ints(self.encrypt_limbs(r, None))


def modexp_grouped_limbs(
    moduli: np.ndarray, exps: np.ndarray, bases: np.ndarray, device: int = 0
) -> np.ndarray:
    """``out[g, k] = bases[g, k] ** exps[g] mod moduli[g]`` on limb arrays: ``moduli`` [G, L] (odd),
    ``exps`` [G, Le], ``bases`` [G, per_group, L]."""
    moduli = np.ascontiguousarray(moduli, dtype=np.uint32)
    exps = np.ascontiguousarray(exps, dtype=np.uint32)
    bases = np.ascontiguousarray(bases, dtype=np.uint32)
    if bases.ndim != 3 or moduli.ndim != 2 or exps.ndim != 2:
        raise ValueError("expected moduli [G, L], exps [G, Le], bases [G, per_group, L]")
    groups, per_group, limbs = bases.shape
    if moduli.shape != (groups, limbs) or exps.shape[0] != groups:
        raise ValueError("shape mismatch between moduli / exps / bases")
    out = np.zeros_like(bases)
    _native.check(
        _native.lib.dkg_modexp_grouped(
            device, moduli.ctypes.data, exps.ctypes.data, exps.shape[1], bases.ctypes.data,
            out.ctypes.data, groups, per_group, limbs,
        )
    )
    return out


def modexp_grouped(
    moduli: Sequence[int], exps: Sequence[int], bases: Sequence[Sequence[int]], device: int = 0
) -> list[list[int]]:
    """Python-int wrapper: group g raises each of ``bases[g]`` to ``exps[g]`` modulo ``moduli[g]``.
    Groups may have different numbers of bases (padded internally)."""
    if not moduli:
        return []
    if any(m <= 0 or m % 2 == 0 for m in moduli):
        raise ValueError("every modulus must be a positive odd integer")
    if any(e < 0 for e in exps):
        raise ValueError("grouped modexp takes non-negative exponents")
    limbs = limbs_for_bits(max(m.bit_length() for m in moduli))
    exp_limbs = limbs_for_bits(max(max(e.bit_length() for e in exps), 1))
    per_group = max(max(len(b) for b in bases), 1)
    flat: list[int] = []
    for m, bs in zip(moduli, bases):
        flat.extend(b % m for b in bs)
        flat.extend([1] * (per_group - len(bs)))
    arr = ints_to_limbs(flat, limbs).reshape(len(moduli), per_group, limbs)
    out = modexp_grouped_limbs(ints_to_limbs(moduli, limbs), ints_to_limbs(exps, exp_limbs), arr, device)
    vals = limbs_to_ints(out.reshape(-1, limbs))
    return [vals[g * per_group : g * per_group + len(bs)] for g, bs in enumerate(bases)]


def biprime_v_batch_limbs(
    moduli: np.ndarray, exps: np.ndarray, gvals: np.ndarray, correct: int, device: int = 0
) -> tuple[np.ndarray, np.ndarray]:
    """Fused v calculation of one ``compute_modulus`` round on limb arrays: Jacobi filter, the
    first ``correct`` usable g's per candidate, grouped modexp.  ``moduli`` [G, L], ``exps`` [G, Le],
    ``gvals`` [G, NG, L].  Returns (v [G, correct, L], count [G])."""
    moduli = np.ascontiguousarray(moduli, dtype=np.uint32)
    exps = np.ascontiguousarray(exps, dtype=np.uint32)
    gvals = np.ascontiguousarray(gvals, dtype=np.uint32)
    groups, ng, limbs = gvals.shape
    if moduli.shape != (groups, limbs) or exps.shape[0] != groups:
        raise ValueError("shape mismatch between moduli / exps / gvals")
    out = np.zeros((groups, correct, limbs), dtype=np.uint32)
    count = np.zeros(groups, dtype=np.int32)
    _native.check(
        _native.lib.dkg_biprime_v_batch(
            device, moduli.ctypes.data, exps.ctypes.data, exps.shape[1], gvals.ctypes.data, ng, correct,
            out.ctypes.data, count.ctypes.data, groups, limbs,
        )
    )
    return out, count


def jacobi_batch(moduli: Sequence[int], gvals: Sequence[Sequence[int]], device: int = 0) -> list[list[int]]:
    """Jacobi symbols (g / N) for every g of every candidate N (odd), on the GPU."""
    if not moduli:
        return []
    if any(m <= 0 or m % 2 == 0 for m in moduli):
        raise ValueError("n should be an odd positive integer")
    limbs = limbs_for_bits(max(m.bit_length() for m in moduli))
    per_group = max(max(len(g) for g in gvals), 1)
    flat: list[int] = []
    for m, gs in zip(moduli, gvals):
        flat.extend(g % m for g in gs)
        flat.extend([0] * (per_group - len(gs)))
    g_arr = ints_to_limbs(flat, limbs)
    m_arr = ints_to_limbs(moduli, limbs)
    sym = np.zeros((len(moduli), per_group), dtype=np.int8)
    _native.check(
        _native.lib.dkg_jacobi_batch(device, m_arr.ctypes.data, g_arr.ctypes.data, per_group, sym.ctypes.data,
                                     len(moduli), limbs)
    )
    return [[int(x) for x in sym[i, : len(gs)]] for i, gs in enumerate(gvals)]


def small_prime_sieve(moduli: Sequence[int], primes: Sequence[int], device: int = 0) -> list[bool]:
    """``True`` where the candidate has a divisor in ``primes`` (all < 2^32)."""
    if not moduli:
        return []
    if not primes:
        return [False] * len(moduli)
    limbs = limbs_for_bits(max(max(m.bit_length() for m in moduli), 1))
    m_arr = ints_to_limbs(moduli, limbs)
    p_arr = np.ascontiguousarray(np.array(list(primes), dtype=np.uint64).astype(np.uint32))
    flags = np.zeros(len(moduli), dtype=np.uint8)
    _native.check(
        _native.lib.dkg_small_prime_sieve(device, m_arr.ctypes.data, p_arr.ctypes.data, len(p_arr), flags.ctypes.data,
                                          len(moduli), limbs)
    )
    return [bool(x) for x in flags]


def biprime_verdict(
    moduli: Sequence[int], v_by_party: dict[int, Sequence[Sequence[int]]], correct: int, device: int = 0
) -> list[bool]:
    """Verdict of the biprimality test for every candidate: ``v_by_party[i][g]`` is party i's v list
    for candidate g (party 1 is the left-hand side).  A candidate with fewer than ``correct`` v
    values from any party fails, as do candidates with a failing test."""
    if not moduli:
        return []
    groups = len(moduli)
    parties = sorted(v_by_party)
    if parties[0] != 1:
        raise KeyError(1)
    limbs = limbs_for_bits(max(m.bit_length() for m in moduli))
    enough = [all(len(v_by_party[p][g]) >= correct for p in parties) for g in range(groups)]
    arr = np.zeros((len(parties), groups, correct, limbs), dtype=np.uint32)
    for pi, p in enumerate(parties):
        flat = []
        for g in range(groups):
            vs = list(v_by_party[p][g][:correct]) if enough[g] else []
            flat.extend(x % moduli[g] for x in vs)
            flat.extend([0] * (correct - len(vs)))
        arr[pi] = ints_to_limbs(flat, limbs).reshape(groups, correct, limbs)
    ok = np.zeros(groups, dtype=np.uint8)
    m_arr = ints_to_limbs(moduli, limbs)
    _native.check(
        _native.lib.dkg_biprime_verdict(device, m_arr.ctypes.data, arr.ctypes.data, len(parties), correct,
                                        ok.ctypes.data, groups, limbs)
    )
    return [bool(o) and e for o, e in zip(ok, enough)]


class ThresholdContext:
    """In-process threshold decryption sharded over the GPUs of one box (C ABI
    ``dkg_threshold_*``): all d+1 parties' exponents for one key, i.e. both loops of
    ``DistributedPaillier._decrypt_sequence_raw`` (``distributed_keygen.py:463-466, 510-515``) for
    every party, with the ciphertexts uploaded once and the partials kept on the device.
    ``exponents`` maps party index (1..d+1) to its signed exponent
    (``PaillierSharedKey.partial_decrypt_exponent``)."""

    def __init__(self, n: int, theta_inv: int, exponents: dict[int, int], devices: Sequence[int] | None = None) -> None:
        parties = sorted(exponents)
        if parties != list(range(1, len(parties) + 1)):
            raise KeyError("exponents of parties 1..d+1 are needed")
        self.n = n
        self.shares = len(parties)
        self.devices = list(devices) if devices is not None else list(range(max(1, _native.device_count())))
        self.n_limbs = limbs_for_bits(n.bit_length())
        self.n2_limbs = limbs_for_bits((n * n).bit_length())
        exp_limbs = limbs_for_bits(max(max(abs(e).bit_length() for e in exponents.values()), 1))
        exps = ints_to_limbs([abs(exponents[p]) for p in parties], exp_limbs)
        neg = np.array([1 if exponents[p] < 0 else 0 for p in parties], dtype=np.uint8)
        dev = np.array(self.devices, dtype=np.int32)
        self._n = int_to_limbs(n, self.n_limbs)
        self._th = int_to_limbs(theta_inv % n, self.n_limbs)
        handle = ctypes.c_void_p()
        _native.check(_native.lib.dkg_threshold_ctx_create(
            dev.ctypes.data, len(self.devices), self._n.ctypes.data, self.n_limbs, self._th.ctypes.data, self.shares,
            exps.ctypes.data, exp_limbs, neg.ctypes.data, ctypes.byref(handle)))
        self._h = handle

    def close(self) -> None:
        if getattr(self, "_h", None):
            _native.lib.dkg_threshold_ctx_destroy(self._h)
            self._h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def _rows(self, arr: np.ndarray, width: int) -> np.ndarray:
        arr = np.ascontiguousarray(arr, dtype=np.uint32)
        if arr.shape[-1] != width:
            raise ValueError(f"rows must have {width} limbs")
        return arr

    def decrypt_limbs(self, ciphertexts: np.ndarray, want_partials: bool = False,
                      out: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray, np.ndarray | None]:
        """[count, n2_limbs] -> (plaintexts [count, n_limbs], status [count], partials [shares, count, n2_limbs] | None)."""
        cts = self._rows(ciphertexts, self.n2_limbs)
        count = cts.shape[0]
        plain = out if out is not None else np.zeros((count, self.n_limbs), dtype=np.uint32)
        status = np.zeros(count, dtype=np.uint8)
        parts = np.zeros((self.shares, count, self.n2_limbs), dtype=np.uint32) if want_partials else None
        _native.check(_native.lib.dkg_threshold_decrypt_batch(
            self._h, cts.ctypes.data, plain.ctypes.data, parts.ctypes.data if parts is not None else None,
            status.ctypes.data, count))
        return plain, status, parts

    def partials_limbs(self, ciphertexts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Every party's partial decryptions of the same ciphertexts in one call (one shared squaring
        chain on the device): [count, n2_limbs] -> (partials [shares, count, n2_limbs], status
        [shares, count])."""
        cts = self._rows(ciphertexts, self.n2_limbs)
        count = cts.shape[0]
        parts = np.zeros((self.shares, count, self.n2_limbs), dtype=np.uint32)
        status = np.zeros((self.shares, count), dtype=np.uint8)
        _native.check(_native.lib.dkg_threshold_partials_batch(
            self._h, cts.ctypes.data, parts.ctypes.data, status.ctypes.data, count))
        return parts, status

    def partial_decrypt_limbs(self, party: int, ciphertexts: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        cts = self._rows(ciphertexts, self.n2_limbs)
        count = cts.shape[0]
        out = np.zeros_like(cts)
        status = np.zeros(count, dtype=np.uint8)
        _native.check(_native.lib.dkg_threshold_partial_decrypt_batch(
            self._h, party - 1, cts.ctypes.data, out.ctypes.data, status.ctypes.data, count))
        return out, status

    def combine_limbs(self, partials: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        parts = self._rows(partials, self.n2_limbs)
        if parts.ndim != 3 or parts.shape[0] != self.shares:
            raise ValueError(f"partials must have shape [{self.shares}, count, {self.n2_limbs}]")
        count = parts.shape[1]
        plain = np.zeros((count, self.n_limbs), dtype=np.uint32)
        status = np.zeros(count, dtype=np.uint8)
        _native.check(_native.lib.dkg_threshold_combine_batch(
            self._h, parts.ctypes.data, plain.ctypes.data, status.ctypes.data, count))
        return plain, status


class pinned:
    """Context manager: page-lock a numpy array for the duration of the block (full-rate PCIe copies)."""

    def __init__(self, *arrays: np.ndarray) -> None:
        self.arrays = [a for a in arrays if a is not None and a.nbytes]

    def __enter__(self):
        done = []
        try:
            for a in self.arrays:
                _native.check(_native.lib.dkg_host_register(a.ctypes.data, a.nbytes))
                done.append(a)
        except Exception:
            for a in done:
                _native.lib.dkg_host_unregister(a.ctypes.data)
            raise
        return self

    def __exit__(self, *exc) -> None:
        for a in self.arrays:
            _native.lib.dkg_host_unregister(a.ctypes.data)
