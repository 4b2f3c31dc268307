"""Input sequence and digest function shared by make_golden.py (which records the reference's
outputs) and the tests (which regenerate the inputs and compare digests)."""
from __future__ import annotations

import hashlib
import random


def digest_inputs(n: int, seed: int, count: int) -> list[tuple[int, int]]:
    """The (m, r) sequence of a digest set: a pure function of (n, seed, count)."""
    rng = random.Random(seed + 77)
    return [(rng.randrange(n), rng.randrange(1, n)) for _ in range(count)]


def sha(v: int) -> str:
    return hashlib.sha256(v.to_bytes((v.bit_length() + 7) // 8 or 1, "big")).hexdigest()
