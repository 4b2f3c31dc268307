"""GPU parity: the CUDA modexp path (through the C ABI) against the oracle / golden vectors."""
from __future__ import annotations

import base64
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


@pytest.fixture(scope="module")
def eng():
    import protocols.distributed_keygen_b200 as eng_mod
    from protocols.distributed_keygen_b200 import _native

    assert _native.device_count() >= 1, "no CUDA device: the gpu tests need a B200"
    return eng_mod


def test_small_moduli_random(eng):
    """Random odd moduli of many widths (every kernel shape), signed exponents, edge bases."""
    rng = random.Random(2026)
    for bits in [33, 64, 67, 96, 134, 200, 256, 515, 768, 1030, 1536, 2048, 2052]:
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        for e in [0, 1, 2, 3, 65537, rng.getrandbits(bits + 60), rng.getrandbits(17)]:
            ctx = eng.ModexpContext(n, e)
            bases = [0, 1, 2, n - 1, n - 2] + [rng.randrange(n) for _ in range(40)]
            got = ctx.modexp(bases)
            assert got == [pow(b, e, n) for b in bases], (bits, e)
            ctx.close()


def test_negative_exponent_and_status(eng):
    import math

    rng = random.Random(7)
    p, q = 1000003, 999983
    for n in [p * q, (p * q) ** 2, rng.getrandbits(300) | 1 | (1 << 299)]:
        e = -rng.getrandbits(100)
        ctx = eng.ModexpContext(n, e)
        bases = [rng.randrange(1, n) for _ in range(70)]
        units = [b for b in bases if math.gcd(b, n) == 1]
        assert ctx.modexp(units) == [pow(b, e, n) for b in units]
        # status flags for non-units (0, multiples of a factor)
        from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

        mixed = units[:5] + [0, p if n % p == 0 else 0, units[5]]
        out, status = ctx.modexp_limbs(ints_to_limbs(mixed, ctx.limbs))
        want_status = [0 if math.gcd(b, n) == 1 else 1 for b in mixed]
        assert list(status) == want_status
        vals = limbs_to_ints(out)
        for b, v, s in zip(mixed, vals, want_status):
            assert v == (0 if s else pow(b, e, n))
        with pytest.raises(ZeroDivisionError):
            ctx.modexp([0])
        ctx.close()


def test_negative_exponent_large_batch_with_non_units(eng):
    """Batched inversion (Montgomery's trick along chains) with non-invertible elements sprinkled
    in: only those elements are flagged, every other result is exact."""
    import math

    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(99)
    p, q = (1 << 127) - 1, (1 << 89) - 1  # Mersenne primes
    n = (p * q) ** 2
    e = -rng.getrandbits(260)
    ctx = eng.ModexpContext(n, e)
    bases = [rng.randrange(1, n) for _ in range(3000)]
    for pos, bad in [(0, 0), (31, p), (32, q * 5), (1500, p * q), (2999, p * p)]:
        bases[pos] = bad % n
    out, status = ctx.modexp_limbs(ints_to_limbs(bases, ctx.limbs))
    vals = limbs_to_ints(out)
    for i, b in enumerate(bases):
        if math.gcd(b, n) != 1:
            assert status[i] == 1 and vals[i] == 0, i
        else:
            assert status[i] == 0 and vals[i] == pow(b, e, n), i
    ctx.close()


def test_square_modulus_pair_arithmetic_matches_direct_kernel(eng):
    """Contexts created with the root N (pair arithmetic modulo N, csrc/dkg_nsq.cuh) against
    contexts on N^2 itself and against pow(): all widths, signed exponents, non-units."""
    import math

    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(404)
    for bits in [9, 20, 61, 67, 130, 257, 515, 1030, 2048, 2051]:
        p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
        q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
        n, n2 = p * q, (p * q) ** 2
        for e in [0, 1, 3, rng.getrandbits(bits + 40), -rng.getrandbits(bits + 40)]:
            fast = eng.ModexpContext(n2, e, root=n)
            assert fast.info()["pair_arithmetic"] == 1
            bases = [0, 1, n2 - 1, n, n + 1, n - 1] + [rng.randrange(n2) for _ in range(70)]
            if e < 0:
                bases = [b for b in bases if math.gcd(b, n) == 1]
            assert fast.modexp(bases) == [pow(b, e, n2) for b in bases], (bits, e)
            fast.close()
    # non-units under a negative exponent: exact per-element status through the fallback
    p, q = (1 << 61) - 1, (1 << 31) - 1
    n, n2 = p * q, (p * q) ** 2
    e = -rng.getrandbits(150)
    fast = eng.ModexpContext(n2, e, root=n)
    bases = [rng.randrange(1, n2) for _ in range(500)]
    bases[3], bases[77], bases[499] = p, 0, q * q
    out, status = fast.modexp_limbs(ints_to_limbs(bases, fast.limbs))
    vals = limbs_to_ints(out)
    for i, b in enumerate(bases):
        if math.gcd(b, n) != 1:
            assert status[i] == 1 and vals[i] == 0
        else:
            assert status[i] == 0 and vals[i] == pow(b, e, n2)
    fast.close()
    with pytest.raises(ValueError):
        eng.ModexpContext(n2, 5, root=n + 2)


def test_partial_decrypt_golden_fixture_keys(eng, fixture_vectors):
    """The reference's 24 golden keys: c^(e_i) mod N^2 must equal what the reference's
    PaillierSharedKey.partial_decrypt returned (tests/golden/fixture_vectors.json)."""
    from oracle import keys as okeys

    for entry in fixture_vectors["sets"]:
        keys = {k["player_id"]: okeys.key_from_blob(base64.b64decode(k["blob_b64"])) for k in entry["keys"]}
        vectors = [v for v in entry["vectors"] if "error" not in v]
        cs = [_h(v["c"]) for v in vectors]
        for pid, key in keys.items():
            ctx = eng.ModexpContext(key.n_square, key.partial_decrypt_exponent())
            got = ctx.modexp(cs)
            assert got == [_h(v["partials"][str(pid)]) for v in vectors]
            ctx.close()


def test_partial_decrypt_golden_dealer_keys(eng, dealer_vectors):
    """Reference-shaped synthetic keys at 128/512/2048/4096 bits (values recorded from the
    reference's partial_decrypt)."""
    from oracle import keys as okeys

    for name, item in dealer_vectors["keys"].items():
        dk = okeys.dealer_key_from_json(item["key"])
        vectors = [v for v in item["vectors"] if "error" not in v]
        cs = [_h(v["c"]) for v in vectors]
        for pid, key in dk.keys.items():
            ctx = eng.ModexpContext(key.n_square, key.partial_decrypt_exponent())
            assert ctx.modexp(cs) == [_h(v["partials"][str(pid)]) for v in vectors], (name, pid)
            ctx.close()


def test_ragged_batch_sizes(eng):
    """Batch sizes around the warp (32) and wave boundaries; empty batch."""
    rng = random.Random(5)
    n = rng.getrandbits(521) | 1 | (1 << 520)
    e = rng.getrandbits(530)
    ctx = eng.ModexpContext(n, e)
    assert ctx.modexp([]) == []
    for count in [1, 31, 32, 33, 63, 65, 1000]:
        bases = [rng.randrange(n) for _ in range(count)]
        assert ctx.modexp(bases) == [pow(b, e, n) for b in bases]
    ctx.close()


@pytest.mark.parametrize("sliding", ["0", "1"])
def test_window_exponent_shapes(eng, sliding, monkeypatch):
    """Exponents that stress the operation list, with the default fixed windows (digit 0 multiplies
    by one) and with sliding windows (DKG_SLIDING_WINDOW=1): powers of two, long zero runs (more
    than 255 squarings in one step), all ones, alternating bits; through the generic kernel and
    through the pair arithmetic (N^2 with root)."""
    monkeypatch.setenv("DKG_SLIDING_WINDOW", sliding)
    rng = random.Random(77)
    p = rng.getrandbits(130) | 1 | (1 << 129)
    q = rng.getrandbits(130) | 1 | (1 << 129)
    root = p * q
    exps = [1 << 300, (1 << 300) + 1, (1 << 521) - 1, (1 << 400) | 1, int("10" * 150, 2), int("1" + "0" * 299 + "1" + "0" * 270 + "111", 2),
            -((1 << 260) + 5)]
    for modulus, kw in ((rng.getrandbits(520) | 1 | (1 << 519), {}), (root * root, {"root": root})):
        bases = [1, 2, modulus - 1] + [rng.randrange(2, modulus) for _ in range(45)]
        for e in exps:
            ctx = eng.ModexpContext(modulus, e, **kw)
            if e < 0:
                import math
                bases_e = [b for b in bases if math.gcd(b, modulus) == 1]
                want = [pow(pow(b, -1, modulus), -e, modulus) for b in bases_e]
            else:
                bases_e, want = bases, [pow(b, e, modulus) for b in bases]
            assert ctx.modexp(bases_e) == want, (modulus.bit_length(), e.bit_length(), kw.keys())
            ctx.close()
