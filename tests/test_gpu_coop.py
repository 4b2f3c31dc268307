"""GPU parity of the cooperative (warp-per-operand) latency kernels, csrc/dkg_coop.cuh, through the
C ABI: the small-batch route of partial decryption (``paillier_shared_key.py:52-93``: one
ciphertext in ``_decrypt_raw``, ten in the reference's sequence test), of the encryption randomness
and of the biprimality-test v calculation (``distributed_keygen.py:1056-1108``).  Checked against
CPython ``pow`` (what the reference's ``pow_mod`` computes), against the values recorded from the
reference's own classes, and against the thread-per-operand kernels on the same inputs."""
from __future__ import annotations

import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


@pytest.fixture()
def native():
    from protocols.distributed_keygen_b200 import _native

    saved = (_native.config_get("coop_max"), _native.config_get("coop_grouped_max"))
    yield _native
    _native.config_set("coop_max", saved[0])
    _native.config_set("coop_grouped_max", saved[1])


def _ctx(native, modulus, exponent, root, coop: bool):
    from protocols.distributed_keygen_b200 import ModexpContext

    native.config_set("coop_max", 8192 if coop else 0)
    return ModexpContext(modulus, exponent, root=root)


@pytest.mark.parametrize("bits", [40, 67, 130, 515, 1030, 2048, 2051, 3000, 4099])
def test_coop_pair_modexp_equals_pow(native, bits):
    rng = random.Random(bits)
    p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
    q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
    n = p * q
    n2 = n * n
    count = 3 if bits > 2100 else 7
    for ebits, sign in ((1, 1), (9, -1), (min(2 * bits + 100, 300 if bits > 2100 else 10**9), 1), (min(bits, 200 if bits > 2100 else 10**9), -1)):
        e = sign * (rng.getrandbits(ebits) | (1 << (ebits - 1)))
        bases = [rng.randrange(1, n2) for _ in range(count)] + [1, n2 - 1, n + 1]
        bases = [b for b in bases if math.gcd(b, n) == 1]
        ctx = _ctx(native, n2, e, n, True)
        assert ctx.info()["pair_arithmetic"] == 1
        got = ctx.modexp(bases)
        ctx.close()
        assert got == [pow(b, e, n2) for b in bases], (bits, ebits, sign)


def test_coop_matches_thread_per_operand_kernel_and_golden(native, dealer_vectors):
    from oracle import keys as okeys

    for name, entry in dealer_vectors["keys"].items():
        dk = okeys.dealer_key_from_json(entry["key"])
        good = [v for v in entry["vectors"] if "error" not in v]
        cs = [_h(v["c"]) for v in good]
        for pid, key in dk.keys.items():
            e = key.partial_decrypt_exponent()
            want = [_h(v["partials"][str(pid)]) for v in good]
            a = _ctx(native, key.n_square, e, key.n, True)
            got_coop = a.modexp(cs)
            a.close()
            assert got_coop == want, (name, pid, "cooperative kernel differs from the reference's partial_decrypt")
            if name.startswith("cfg1") or name.startswith("small"):
                b = _ctx(native, key.n_square, e, key.n, False)
                assert b.modexp(cs) == want
                b.close()


@pytest.mark.parametrize("coop", [True, False])
def test_negative_exponent_status_per_element(native, coop):
    """Non-units under a negative exponent: status 1 on exactly those rows, zero rows, the others
    correct (the reference raises ZeroDivisionError from mod_inv per ciphertext).  Cooperative route:
    per-instance inversion.  Thread-per-operand route: the batched inversion flags the chain on the
    device and the predicated direct kernel redoes the call -- no host round trip in between."""
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(11)
    p, q = 1000003, 999983
    n = p * q
    n2 = n * n
    e = -(rng.getrandbits(90) | 1)
    bases = [rng.randrange(1, n2) for _ in range(200)]
    bad = {3: p * 12345, 7: q * q * 5, 11: 0, 19: n, 150: p, 199: q * 77}
    for i, v in bad.items():
        bases[i] = v % n2
    ctx = _ctx(native, n2, e, n, coop)
    out, status = ctx.modexp_limbs(ints_to_limbs(bases, ctx.limbs))
    ctx.close()
    vals = limbs_to_ints(out)
    for i, b in enumerate(bases):
        if math.gcd(b, n) != 1:
            assert status[i] == 1 and vals[i] == 0, i
        else:
            assert status[i] == 0 and vals[i] == pow(b, e, n2), i
    with pytest.raises(ZeroDivisionError):
        c2 = _ctx(native, n2, e, n, coop)
        try:
            c2.modexp(bases)
        finally:
            c2.close()


def test_coop_encrypt_small_batches(native, fixture_vectors):
    from protocols.distributed_keygen_b200 import EncryptContext

    native.config_set("coop_max", 8192)
    rng = random.Random(3)
    for n in (_h(fixture_vectors["sets"][0]["keys"][0]["n"]), (rng.getrandbits(1026) | 1 | 1 << 1025) * (rng.getrandbits(1025) | 1 | 1 << 1024)):
        enc = EncryptContext(n)
        ms = [rng.randrange(n) for _ in range(5)] + [0, n - 1]
        rs = [rng.randrange(1, n) for _ in ms]
        assert enc.encrypt(ms, rs) == [(1 + m * n) * pow(r, n, n * n) % (n * n) for m, r in zip(ms, rs)]
        assert enc.randomness(rs[:2]) == [pow(r, n, n * n) for r in rs[:2]]
        enc.close()


@pytest.mark.parametrize("coop", [True, False])
def test_grouped_small_batches_both_kernels(native, coop):
    from protocols.distributed_keygen_b200 import modexp_grouped

    native.config_set("coop_grouped_max", 16384 if coop else 0)
    rng = random.Random(17 + coop)
    for bits, groups, per in ((31, 3, 2), (64, 2, 5), (190, 4, 3), (1024, 3, 4), (2052, 5, 8), (2052, 1, 1), (2080, 2, 3)):
        moduli = [rng.getrandbits(bits) | 1 | (1 << (bits - 1)) for _ in range(groups)]
        moduli[0] = moduli[0] if bits > 31 else 1   # N = 1: everything is 0
        exps = [rng.getrandbits(rng.choice([1, 7, bits // 2, bits - 2])) for _ in range(groups)]
        exps[-1] = 0
        bases = [[rng.randrange(m) for _ in range(per)] for m in moduli]
        if bits > 31:
            bases[0][0] = 0
            bases[-1][-1] = moduli[-1] - 1
        got = modexp_grouped(moduli, exps, bases)
        want = [[pow(b, e, m) for b in bs] for m, e, bs in zip(moduli, exps, bases)]
        assert got == want, (bits, coop)


def test_biprime_v_small_round_matches_reference_values(native, biprime_vectors):
    """One compute_modulus round's v calculation at the reference's real sizes (a handful of
    candidates): the cooperative grouped kernel against the v values the reference recorded."""
    from protocols.distributed_keygen_b200 import distributed_keygen as dk

    native.config_set("coop_grouped_max", 16384)
    for case in biprime_vectors["cases"]:
        n = _h(case["n"])
        g = [_h(x) for x in case["g_values"]]
        for i in range(1, case["parties"] + 1):
            want = [_h(v) for v in case["v"][str(i)]]
            got = dk.biprime_test_v_calculation_batch(
                [(g, n, _h(case["p_shares"][i - 1]), _h(case["q_shares"][i - 1]))] * 2, i, case["correct_param_biprime"])
            assert got == [want, want], (case["key_length"], i)


def test_coop_and_wave_kernels_agree_on_a_mixed_batch(native):
    """The same 300 ciphertexts through both routes (2048-bit reference-shaped modulus, negative
    exponent): bit-identical rows."""
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs

    rng = random.Random(23)
    p = rng.getrandbits(1026) | 1 | (1 << 1025)
    q = rng.getrandbits(1026) | 1 | (1 << 1025)
    n = p * q
    n2 = n * n
    e = -(rng.getrandbits(4200) | 1)
    rows = ints_to_limbs([rng.randrange(1, n2) for _ in range(300)], (n2.bit_length() + 31) // 32)
    a = _ctx(native, n2, e, n, True)
    out_a, st_a = a.modexp_limbs(rows)
    a.close()
    b = _ctx(native, n2, e, n, False)
    out_b, st_b = b.modexp_limbs(rows)
    b.close()
    # p, q are random odd numbers: many rows share a factor with n and must be flagged, identically
    from protocols.distributed_keygen_b200.limbs import limbs_to_ints

    vals = limbs_to_ints(rows)
    unit = np.array([math.gcd(v, n) == 1 for v in vals])
    assert unit.any() and (~unit).any()
    assert np.array_equal(st_a == 0, unit) and np.array_equal(st_b == 0, unit)
    assert np.array_equal(out_a, out_b)
    # spot check against CPython
    idx = [int(i) for i in np.flatnonzero(unit)[:4]]
    assert [limbs_to_ints(out_a[i : i + 1])[0] for i in idx] == [pow(vals[i], e, n2) for i in idx]


@pytest.mark.parametrize("coop", [True, False])
def test_rows_not_below_the_modulus_are_flagged(native, coop):
    """The limb API takes values < modulus; a row that fits the limb width but is >= N^2 must come
    back with status 3 and a zero row, never as a silently wrong residue (both routes)."""
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(31)
    n = 1000003 * 999983
    n2 = n * n
    top = 1 << (32 * ((n2.bit_length() + 31) // 32))
    for e in (65537, -65537):
        vals = [rng.randrange(1, n2) for _ in range(40)]
        vals = [v for v in vals if math.gcd(v, n) == 1]
        bad = {2: n2, 5: n2 + 1, 9: top - 1}
        for i, v in bad.items():
            vals[i] = v
        ctx = _ctx(native, n2, e, n, coop)
        out, st = ctx.modexp_limbs(ints_to_limbs(vals, ctx.limbs))
        ctx.close()
        got = limbs_to_ints(out)
        for i, v in enumerate(vals):
            if i in bad:
                assert st[i] == 3 and got[i] == 0
            else:
                assert st[i] == 0 and got[i] == pow(v, e, n2)
    # generic odd modulus (direct kernel)
    from protocols.distributed_keygen_b200 import ModexpContext

    m = rng.getrandbits(200) | 1 | (1 << 199)
    ctx = ModexpContext(m, 12345)
    vals = [rng.randrange(m) for _ in range(10)] + [m, m + 7]
    out, st = ctx.modexp_limbs(ints_to_limbs(vals, ctx.limbs))
    ctx.close()
    assert list(st) == [0] * 10 + [3, 3]
    assert limbs_to_ints(out)[:10] == [pow(v, 12345, m) for v in vals[:10]]


@pytest.mark.parametrize("coop", [True, False])
def test_constant_time_table_access_mode(native, coop):
    """config "ct_table" (env DKG_CT_TABLE=1): every multiplication scans the whole window table
    under a mask instead of indexing it with the exponent digit; same results, narrower windows."""
    rng = random.Random(41)
    native.config_set("ct_table", 1)
    try:
        for bits in (67, 515, 2051):
            p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
            q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
            n = p * q
            n2 = n * n
            for sign in (1, -1):
                ebits = 2 * bits + 60 if bits < 2000 else 700
                e = sign * (rng.getrandbits(ebits) | (1 << (ebits - 1)))
                # digits 0 inside the exponent exercise the masked Montgomery one
                e = sign * (abs(e) & ~(0xFFFFF << 40))
                bases = [b for b in (rng.randrange(1, n2) for _ in range(12)) if math.gcd(b, n) == 1]
                ctx = _ctx(native, n2, e, n, coop)
                info = ctx.info()
                assert info["pair_arithmetic"] == 1 and info["window_bits"] <= 5
                assert ctx.modexp(bases) == [pow(b, e, n2) for b in bases], (bits, sign, coop)
                ctx.close()
    finally:
        native.config_set("ct_table", 0)


def test_cooperative_combine_equals_word_serial_kernel_and_python(native, monkeypatch):
    """Share combination (paillier_shared_key.py:108-125): the warp-per-ciphertext kernel against the
    thread-per-ciphertext one (DKG_COOP_COMBINE=0) and against Python integers, including the edge
    values of the divisibility check: product 0, product 1 (message 0), tampered partials."""
    from protocols.distributed_keygen_b200 import CombineContext
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    rng = random.Random(5)
    for bits, shares in ((67, 3), (515, 3), (2051, 5), (2048, 2), (3000, 3)):
        p = rng.getrandbits(bits // 2) | 1 | (1 << (bits // 2 - 1))
        q = rng.getrandbits(bits - bits // 2) | 1 | (1 << (bits - bits // 2 - 1))
        n = p * q
        n2 = n * n
        theta_inv = rng.randrange(1, n)
        count = 40
        cols = []
        for i in range(count):
            parts = [rng.randrange(1, n2) for _ in range(shares - 1)]
            prod = 1
            for v in parts:
                prod = prod * v % n2
            # last partial chosen so that the product is 1 + m N (divisible), except every 5th
            m = rng.randrange(n)
            target = (1 + m * n) % n2 if i % 5 else rng.randrange(n2)
            if i == 7:
                target = 1          # message 0: y = 0
            try:
                last = target * pow(prod, -1, n2) % n2
            except ValueError:
                last = rng.randrange(n2)
            if i == 9:
                last = 0            # product 0
            cols.append(parts + [last])
        arr = np.stack([ints_to_limbs([cols[i][s] for i in range(count)], (n2.bit_length() + 31) // 32) for s in range(shares)])
        want_val, want_st = [], []
        for i in range(count):
            x = 1
            for v in cols[i]:
                x = x * v % n2
            ok = (x - 1) % n == 0
            want_st.append(0 if ok else 2)
            want_val.append(((x - 1) // n * theta_inv) % n if ok else 0)
        outs = []
        for coop in ("1", "0"):
            monkeypatch.setenv("DKG_COOP_COMBINE", coop)
            ctx = CombineContext(n, theta_inv, shares)
            out, st = ctx.combine_limbs(arr)
            ctx.close()
            assert list(st) == want_st, (bits, coop)
            assert limbs_to_ints(out) == want_val, (bits, coop)
            outs.append(out)
        assert np.array_equal(outs[0], outs[1])
