"""Integer model of the shared squaring chain (csrc/dkg_nsq.cuh ``modexp_nsq_multi_kernel``, host side
``setup_threshold_multi`` in csrc/dkg_engine.cu): right-to-left digits, one dummy bucket for digit 0,
running-product fold of the buckets.  The model mirrors the kernel's loop structure statement by
statement; the GPU kernel itself is compared with the reference's values in
tests/test_gpu_shared_squarings.py."""
from __future__ import annotations

import random

import pytest


def digits_lsb_first(e: int, w: int, nwin: int) -> list[int]:
    return [(e >> (k * w)) & ((1 << w) - 1) for k in range(nwin)]


def shared_chain(c: int, exponents: list[int], w: int, mod: int) -> tuple[list[int], dict]:
    """All c^e mod `mod` for the non-negative `exponents` with ONE chain of squarings."""
    D = 1 << w
    ebits = max(max(e.bit_length() for e in exponents), 1)
    nwin = (ebits + w - 1) // w
    digs = [digits_lsb_first(e, w, nwin) for e in exponents]
    buckets = [[1] * D for _ in exponents]
    count = {"sqr": 0, "mul": 0}
    cur = c % mod
    for k in range(nwin):
        if k > 0:
            for _ in range(w):
                cur = cur * cur % mod
                count["sqr"] += 1
        for q in range(len(exponents)):
            d = digs[q][k]                      # digit 0: the dummy bucket, never read again
            buckets[q][d] = buckets[q][d] * cur % mod
            count["mul"] += 1
    out = []
    for q in range(len(exponents)):
        t = buckets[q][D - 1]
        s = t
        for d in range(D - 2, 0, -1):
            t = t * buckets[q][d] % mod
            s = s * t % mod
            count["mul"] += 2
        out.append(s)
    return out, count


@pytest.mark.parametrize("w", [1, 2, 3, 4, 5, 6, 7, 8])
def test_bucket_method_matches_pow(w):
    rng = random.Random(w)
    mod = rng.getrandbits(256) | (1 << 255) | 1
    for parties in (1, 2, 3, 5):
        exps = [rng.getrandbits(rng.choice([1, 7, 64, 200])) for _ in range(parties)]
        exps[0] = 0 if parties > 2 else exps[0]
        c = rng.randrange(2, mod)
        got, _ = shared_chain(c, exps, w, mod)
        assert got == [pow(c, e, mod) for e in exps]


def test_operation_counts_of_the_headline_shape():
    """4190-bit exponents, w = 6, 3 parties: 4182 squarings and 3 * (699 + 124) multiplications, the
    numbers DESIGN.md and bench.py's executed-work formula use."""
    rng = random.Random(1)
    exps = [rng.getrandbits(4190) | (1 << 4189) for _ in range(3)]
    _, count = shared_chain(3, exps, 6, (1 << 127) - 1)
    assert count == {"sqr": 6 * 698, "mul": 3 * (699 + 2 * 62)}
