"""Size-independent properties of the CUDA path on large batches (where comparing every element
with the CPU oracle would take minutes): multiplicativity of the modexp, permutation invariance,
inverse pairs under negated exponents, decrypt(encrypt(m)) == m, additive homomorphism."""
from __future__ import annotations

import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_modexp_multiplicative_and_permutation_invariant_large_batch():
    import protocols.distributed_keygen_b200 as eng
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(12)
    n = (rng.getrandbits(1030) | 1 | (1 << 1029))  # 33-limb modulus (key_length 512 shape)
    e = rng.getrandbits(1100)
    B = 120_000  # several waves, ragged tail
    ctx = eng.ModexpContext(n, e)
    a = [rng.randrange(n) for _ in range(B)]
    b = a[1:] + a[:1]
    ab = [x * y % n for x, y in zip(a, b)]
    ra = ctx.modexp(a)
    rab = ctx.modexp(ab)
    rb = ra[1:] + ra[:1]  # permutation invariance gives modexp(b) for free ...
    assert all(x * y % n == z for x, y, z in zip(ra, rb, rab))  # ... and multiplicativity ties all three
    # a seeded sample against CPython pow
    for i in rng.sample(range(B), 64):
        assert ra[i] == pow(a[i], e, n)
    # shuffled batch == shuffled results
    perm = list(range(B))
    rng.shuffle(perm)
    shuffled = ctx.modexp([a[i] for i in perm[:20000]])
    assert shuffled == [ra[i] for i in perm[:20000]]
    ctx.close()
    # negative exponent: result is the inverse of the positive one
    ctxn = eng.ModexpContext(n, -e)
    import math

    units = [x for x in a[:40000] if math.gcd(x, n) == 1]
    rn = ctxn.modexp(units)
    pos = {x: r for x, r in zip(a, ra)}
    assert all(r * pos[x] % n == 1 for x, r in zip(units, rn))
    ctxn.close()


def test_threshold_paillier_homomorphism_large_batch(dealer_vectors):
    """Enc(m1) * Enc(m2) decrypts to m1 + m2 through encrypt -> 3 partial decryptions -> combine,
    all on the GPU, for a batch spanning several waves (key_length 512, reference-shaped key)."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg1_k512_p3_t1"]["key"])
    n, n2 = dk.n, dk.n * dk.n
    rng = random.Random(31)
    B = 60_000
    m1 = [rng.randrange(-(2**62), 2**62) for _ in range(B)]
    m2 = [rng.randrange(-(2**62), 2**62) for _ in range(B)]
    enc = eng.EncryptContext(n)
    c1 = enc.encrypt(m1, [rng.randrange(1, n) for _ in range(B)])
    c2 = enc.encrypt(m2, [rng.randrange(1, n) for _ in range(B)])
    csum = [x * y % n2 for x, y in zip(c1, c2)]
    keys = {}
    for pid, k in dk.keys.items():
        share = eng.IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
        keys[pid] = eng.PaillierSharedKey(k.n, k.t, pid, share, k.theta)
    got = dkg.decrypt_sequence_local(keys, csum)
    assert got == [(x + y) % n for x, y in zip(m1, m2)]
    enc.close()
    for key in keys.values():
        key.close()
