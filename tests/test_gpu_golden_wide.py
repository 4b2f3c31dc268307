"""Wider golden parity for the hot kernel shapes (more than one warp per shape, a full
wave plus a ragged tail, encryption pinned on reference-validated ciphertexts).

* ``dealer_digests.json``: 64-160 vectors per key (key_length 512 / 2048 exact and reference-shaped /
  2048 with 5 parties t=2 / 4096), inputs regenerated from the seed, the REFERENCE's
  ``partial_decrypt`` / ``decrypt`` outputs recorded as SHA-256 digests by tests/golden/make_golden.py.
  Both routes (cooperative warp-per-operand and thread-per-operand kernels) must reproduce them.
* every ``(m, r, c)`` the reference's ``decrypt`` accepted (``fixture_vectors.json``,
  ``dealer_vectors.json``) must come out of the GPU ``encrypt`` bit for bit (pins a8).
* one launch of the headline shape with a full wave plus 17 rows, 100 % of the rows compared with
  GMP ``mpz_powm`` (what ``gmpy2.powmod`` wraps).
"""
from __future__ import annotations

import base64
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from digests import digest_inputs, sha  # noqa: E402

from conftest import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


@pytest.fixture()
def native():
    from protocols.distributed_keygen_b200 import _native

    saved = _native.config_get("coop_max")
    yield _native
    _native.config_set("coop_max", saved)


@pytest.mark.parametrize("name", ["cfg1_k512_p3_t1", "cfg2_k2048_p3_t1_exact", "cfg2_k2048_p3_t1_real",
                                  "cfg3_k2048_p5_t2_exact", "cfg4_k4096_p3_t1_exact"])
def test_reference_digests_both_routes(native, dealer_vectors, name):
    import protocols.distributed_keygen_b200 as eng
    from oracle import keys as okeys
    from oracle.paillier_oracle import encrypt_raw

    item = load_golden("dealer_digests.json")["keys"][name]
    dk = okeys.dealer_key_from_json(dealer_vectors["keys"][name]["key"])
    assert dk.n == _h(item["n"])
    rows = item["vectors"]
    inputs = digest_inputs(dk.n, item["seed"], len(rows))
    cs = [encrypt_raw(dk.n, m, r) for m, r in inputs]
    assert [sha(c) for c in cs] == [v["c"] for v in rows], "input regeneration out of sync with make_golden.py"
    # a8: the GPU encryption must produce these very ciphertexts
    enc = eng.EncryptContext(dk.n)
    assert enc.encrypt([m for m, _ in inputs], [r for _, r in inputs]) == cs
    enc.close()
    partials = {}
    for coop in (True, False):
        native.config_set("coop_max", 1 << 20 if coop else 0)
        for pid, key in dk.keys.items():
            ctx = eng.ModexpContext(key.n_square, key.partial_decrypt_exponent(), root=key.n)
            got = ctx.modexp(cs)
            ctx.close()
            assert [sha(v) for v in got] == [v["partials"][str(pid)] for v in rows], (name, pid, "coop" if coop else "wave")
            partials[pid] = got
    k1 = dk.keys[1]
    share = eng.IntegerShares(dict(k1.share.shares), k1.share.degree, k1.share.scaling, k1.share.number_of_parties)
    gk = eng.PaillierSharedKey(k1.n, k1.t, 1, share, k1.theta)
    plain = gk.decrypt_batch([{pid: partials[pid][i] for pid in dk.keys} for i in range(len(rows))])
    gk.close()
    assert [sha(v) for v in plain] == [v["plaintext"] for v in rows]
    assert plain == [m for m, _ in inputs]


def test_encrypt_pinned_on_reference_validated_ciphertexts(fixture_vectors, dealer_vectors):
    """Every recorded (m, r, c): the reference's own ``decrypt`` returned m for c (asserted when the
    vectors were made), so c is the value the reference's scheme works with; GPU encrypt(m, r) == c."""
    import protocols.distributed_keygen_b200 as eng

    sets = [(_h(s["keys"][0]["n"]), s["vectors"]) for s in fixture_vectors["sets"]]
    sets += [(_h(k["key"]["n"]), k["vectors"]) for k in dealer_vectors["keys"].values()]
    checked = 0
    for n, vectors in sets:
        good = [v for v in vectors if "m" in v]
        enc = eng.EncryptContext(n)
        got = enc.encrypt([_h(v["m"]) for v in good], [_h(v["r"]) for v in good])
        rn = enc.randomness([_h(v["r"]) for v in good])
        enc.close()
        assert got == [_h(v["c"]) for v in good]
        # r^N itself: c * (1 + m N)^-1 = c * (1 - m N) mod N^2
        assert rn == [_h(v["c"]) * (1 - _h(v["m"]) * n) % (n * n) for v in good]
        checked += len(good)
    assert checked >= 6 * 16 + 20


def test_full_wave_plus_ragged_tail_against_gmp(dealer_vectors):
    """Headline shape (exact 2048-bit key, thread-per-operand pair kernel <14,5>): one launch of a
    full wave + 17 rows -- multi-wave ticketing, the ragged last group -- with EVERY row compared
    with GMP on all host cores."""
    import protocols.distributed_keygen_b200 as eng
    from oracle import gmp as ogmp
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs

    dk = okeys.dealer_key_from_json(dealer_vectors["keys"]["cfg2_k2048_p3_t1_exact"]["key"])
    exps = {pid: k.partial_decrypt_exponent() for pid, k in dk.keys.items()}
    pid = next((p for p, e in exps.items() if e < 0), 1)   # negative: batched inversion + modexp
    e = exps[pid]
    ctx = eng.ModexpContext(dk.n * dk.n, e, root=dk.n)
    info = ctx.info()
    assert info["pair_arithmetic"] == 1 and (info["pair_K"], info["pair_M"]) == (13, 5)
    wave = info["ctas"] * info["warps_per_cta"] * 32
    count = wave + 17
    L = ctx.limbs
    rng = np.random.default_rng(20261017)
    rows = rng.integers(0, 2**32, size=(count, L), dtype=np.uint32)
    rows[:, -1] &= np.uint32(0x3FFFFFFF)       # below N^2 (top limb of an exact 4096-bit N^2 is >= 2^30)
    rows[:, 0] |= np.uint32(1)
    out, status = ctx.modexp_limbs(rows)
    ctx.close()
    assert not status.any()
    mag = abs(e)
    want, _ = ogmp.powm_batch_threads(rows, ints_to_limbs([dk.n * dk.n], L)[0],
                                      ints_to_limbs([mag], (mag.bit_length() + 31) // 32)[0], e < 0, os.cpu_count() or 1)
    bad = np.flatnonzero((out != want).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {count} rows differ from GMP, first: {bad[:5]}"
