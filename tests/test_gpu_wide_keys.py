"""Wide keys (key_length 3072 / 4096 and the widest pair shape): the pair kernels keep the b component
of the running pair in global scratch (``modexp_nsq_kernel<K, M, true>``, csrc/dkg_nsq.cuh) so that
12 warps fit an SM instead of 6-9.  Same values as with both components in shared memory
(``DKG_NSQ_BG=0``) and as CPython ``pow``, for both exponent signs, through the thread-per-ciphertext
route (``ref: paillier_shared_key.py:52-93``; BASELINE.json config 4)."""
from __future__ import annotations

import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def wave_route():
    from protocols.distributed_keygen_b200 import _native

    saved = _native.config_get("coop_max")
    _native.config_set("coop_max", 0)
    yield _native
    _native.config_set("coop_max", saved)


@pytest.mark.parametrize("bits,shape", [(3060, (16, 6)), (3072, (14, 7)), (4096, (12, 11)), (4300, (16, 9))])
def test_b_component_in_global_scratch(wave_route, monkeypatch, bits, shape):
    import protocols.distributed_keygen_b200 as eng

    rng = random.Random(bits)
    n = rng.getrandbits(bits) | (1 << (bits - 1)) | 1
    n2 = n * n
    cs = []
    while len(cs) < 77:      # units only (n is a random odd number, not an RSA modulus)
        c = rng.randrange(1, n2)
        if math.gcd(c, n) == 1:
            cs.append(c)
    cs[3] = 1
    cs[4] = n2 - 1
    for sign in (1, -1):
        e = sign * (rng.getrandbits(300) | (1 << 299))      # short exponent: the arithmetic is what is tested
        want = [pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2) for c in cs]
        warps = {}
        for bg in ("1", "0"):
            monkeypatch.setenv("DKG_NSQ_BG", bg)
            ctx = eng.ModexpContext(n2, e, root=n)
            info = ctx.info()
            assert info["pair_arithmetic"] == 1 and (info["pair_K"], info["pair_M"]) == shape
            warps[bg] = info["pair_warps_per_cta"]
            got = ctx.modexp(cs)
            ctx.close()
            assert got == want, (bits, sign, bg)
        assert warps["1"] == 12 and warps["0"] < 10


def test_full_exponent_4096_bit_key_matches_cpython(wave_route, dealer_vectors):
    """The cfg 4 key with its real 8200-bit exponents, 40 random units per party."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg4_k4096_p3_t1_exact"]["key"])
    n2 = dk.n * dk.n
    rng = random.Random(4)
    cs = [rng.randrange(1, n2) for _ in range(40)]
    for pid, key in dk.keys.items():
        e = key.partial_decrypt_exponent()
        ctx = eng.ModexpContext(n2, e, root=dk.n)
        assert ctx.info()["pair_warps_per_cta"] == 12
        got = ctx.modexp(cs)
        ctx.close()
        assert got == [pow(pow(c, -1, n2), -e, n2) if e < 0 else pow(c, e, n2) for c in cs], pid
