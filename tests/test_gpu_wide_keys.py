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


def test_in_kernel_inversion_option(wave_route, monkeypatch, dealer_vectors):
    """``DKG_INKERNEL_INVERSE=1``: negative exponents inverted inside the pair kernels (pair_invert: GCD on
    the a component + one Newton step in the pair domain) instead of by the batched inversion kernel:
    same partials, exact per-element status for non-units, single-party and shared-chain kernels."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    monkeypatch.setenv("DKG_INKERNEL_INVERSE", "1")
    for name in ("cfg1_k512_p3_t1", "cfg2_k2048_p3_t1_real"):
        dk = okeys.dealer_key_from_json(dealer_vectors["keys"][name]["key"])
        n2 = dk.n * dk.n
        rng = random.Random(len(name))
        cs = [rng.randrange(1, n2) for _ in range(70)]
        cs[9] = sum(dk.p_shares) * 31337 % n2          # not a unit
        rows = ints_to_limbs(cs, (n2.bit_length() + 31) // 32)
        exps = {pid: k.partial_decrypt_exponent() for pid, k in dk.keys.items()}
        assert any(e < 0 for e in exps.values())
        for pid, e in exps.items():
            ctx = eng.ModexpContext(n2, e, root=dk.n)
            out, status = ctx.modexp_limbs(rows)
            ctx.close()
            if e < 0:
                assert status[9] == 1 and not out[9].any() and not np.delete(status, 9).any()
            else:
                assert not status.any()
            keep = [i for i in range(len(cs)) if not (e < 0 and i == 9)]
            assert [limbs_to_ints(out[i : i + 1])[0] for i in keep] == [pow(cs[i], e, n2) for i in keep], (name, pid)
        keys = {}
        for pid, k in dk.keys.items():
            share = eng.IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
            keys[pid] = eng.PaillierSharedKey(k.n, k.t, pid, share, k.theta)
        tctx = dkg.threshold_context(keys, [0])
        parts, status = tctx.partials_limbs(rows)
        tctx.close()
        for pid, e in exps.items():
            keep = [i for i in range(len(cs)) if not (e < 0 and i == 9)]
            assert status[pid - 1, 9] == (1 if e < 0 else 0)
            assert [limbs_to_ints(parts[pid - 1, i : i + 1])[0] for i in keep] == [pow(cs[i], e, n2) for i in keep], (name, pid)
        for k in keys.values():
            k.close()
