"""Host build of the kernels' block-Montgomery templates (csrc/dkg_mont.cuh, carry primitives
emulated in C) against Python big integers: index logic, carry handling, in-place squaring, the
[0, R) invariant and canonical reduction, for every (K, M) shape the engine instantiates."""
from __future__ import annotations

import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "protocols", "distributed_keygen_b200", "csrc")

SHAPES = [(4, 1), (4, 3), (4, 2), (4, 5), (6, 3), (8, 4), (12, 3), (16, 2), (16, 8), (12, 11),
          (22, 3), (22, 6), (16, 16), (14, 5), (12, 6), (14, 7),
          (13, 5), (13, 10), (5, 3), (7, 2), (9, 1)]   # odd block sizes: slots of K + 1 limbs on the device


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("mont_host") / "libmont_host.so"
    subprocess.run(
        ["g++", "-O1", "-w", "-shared", "-fPIC", "-I", CSRC, "-o", str(out),
         os.path.join(ROOT, "tests", "host", "mont_host.cpp")],
        check=True,
    )
    lib = ctypes.CDLL(str(out))
    vp = ctypes.c_void_p
    lib.host_mont.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int]
    lib.host_mont2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp, vp]
    return lib


def _limbs(v: int, n: int) -> np.ndarray:
    return np.frombuffer(v.to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def _val(a: np.ndarray) -> int:
    return int.from_bytes(a.tobytes(), "little")


def _call(lib, K, M, mode, x, y, n, ninv, canon):
    L = K * M
    xa, ya, na, ia = _limbs(x, L), _limbs(y, L), _limbs(n, L), _limbs(ninv, K)
    rc = lib.host_mont(K, M, mode, xa.ctypes.data, ya.ctypes.data, na.ctypes.data, ia.ctypes.data, canon)
    assert rc == 0
    return _val(xa)


@pytest.mark.parametrize("K,M", SHAPES)
def test_mont_mul_host(lib, K, M):
    rng = random.Random(1000 * K + M)
    L = K * M
    R = 1 << (32 * L)
    W = 1 << (32 * K)
    trials = 24 if L <= 64 else 6
    for trial in range(trials):
        bits = rng.choice([32 * L, 32 * L - 1, 32 * L - 40, 32 * L - 63, max(3, 32 * L - 32 * K - 5)])
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        if trial == 0:
            n = R - 1
        ninv = (-pow(n, -1, W)) % W
        r_inv = pow(R, -1, n)
        for mode in (0, 1, 2, 3):
            x, y = rng.randrange(R), rng.randrange(R)
            if trial == 1:
                x = y = R - 1
            if trial == 2:
                x = 0
            want = {0: x * y * r_inv, 1: x * x * r_inv, 2: x * r_inv, 3: x * x * r_inv}[mode] % n
            got = _call(lib, K, M, mode, x, y, n, ninv, 0)
            assert got < R and got % n == want, (K, M, mode, trial)
            # canonical residue after conditional subtraction(s)
            rounds = 1 if mode == 2 else max(1, (R // n).bit_length() + 1)
            if rounds <= 8:
                assert _call(lib, K, M, mode, x, y, n, ninv, rounds) == want, (K, M, mode, trial)


def test_mont_exponentiation_chain_host(lib):
    """A square-and-multiply ladder through the host templates equals pow()."""
    K, M = 4, 5
    L = K * M
    R = 1 << (32 * L)
    W = 1 << (32 * K)
    rng = random.Random(77)
    n = rng.getrandbits(32 * L - 3) | 1 | (1 << (32 * L - 4))
    ninv = (-pow(n, -1, W)) % W
    base, e = rng.randrange(n), rng.getrandbits(200)
    acc = R % n
    b_m = base * R % n
    for bit in bin(e)[2:]:
        acc = _call(lib, K, M, 3, acc, 0, n, ninv, 0)
        if bit == "1":
            acc = _call(lib, K, M, 0, acc, b_m, n, ninv, 0)
    assert _call(lib, K, M, 2, acc, 0, n, ninv, 1) == pow(base, e, n)


@pytest.mark.parametrize("K,M", [(4, 2), (12, 6), (14, 5), (16, 2), (22, 3), (13, 5), (13, 10), (5, 3), (14, 7)])
def test_pair_modes_host(lib, K, M):
    """MONT_MUL2S (x <- 2xy/R) and MONT_MULADD (x <- (xy + s y2)/R) used by the mod N^2 pair
    arithmetic, with operands up to R and a modulus with >= 3 spare bits."""
    rng = random.Random(7000 + 10 * K + M)
    L = K * M
    R = 1 << (32 * L)
    W = 1 << (32 * K)
    for trial in range(12):
        n = rng.getrandbits(32 * L - 3) | 1 | (1 << (32 * L - 4))
        ninv = (-pow(n, -1, W)) % W
        r_inv = pow(R, -1, n)
        x, y = rng.randrange(R), rng.randrange(2 * n)
        s_op, y2 = rng.randrange(2 * n), rng.randrange(R)
        if trial == 0:
            x, y = R - 1, 2 * n - 1
        xa, ya, na, ia = _limbs(x, L), _limbs(y, L), _limbs(n, L), _limbs(ninv, K)
        sa, y2a = _limbs(s_op, L), _limbs(y2, L)
        assert lib.host_mont2(K, M, 4, xa.ctypes.data, ya.ctypes.data, na.ctypes.data, ia.ctypes.data, 0,
                              sa.ctypes.data, y2a.ctypes.data) == 0
        got = _val(xa)
        assert got < R and got % n == (2 * x * y * r_inv) % n, (K, M, trial, "mul2s")
        xa = _limbs(x, L)
        assert lib.host_mont2(K, M, 5, xa.ctypes.data, ya.ctypes.data, na.ctypes.data, ia.ctypes.data, 0,
                              sa.ctypes.data, y2a.ctypes.data) == 0
        got = _val(xa)
        assert got < R and got % n == ((x * y + s_op * y2) * r_inv) % n, (K, M, trial, "muladd")

