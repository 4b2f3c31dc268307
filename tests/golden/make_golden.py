#!/usr/bin/env python
"""
Generate the golden vectors under ``tests/golden/`` by running the REFERENCE's own classes.

Run on the build container only (needs ``/root/reference``; the GPU box never runs this):

    python tests/golden/make_golden.py

What runs reference code and what does not:

* ``PaillierSharedKey.partial_decrypt`` / ``.decrypt`` (``paillier_shared_key.py:52-127``),
  ``DistributedPaillier.__biprime_test_v_calculation`` / ``__biprime_test_with_v_i``
  (``distributed_keygen.py:1056-1175``) and ``utils.mult_list`` are imported from
  ``/root/reference/src`` and executed unmodified.
* Their un-vendored third-party imports are satisfied by the data-holder shims in
  ``tests/golden/ref_shims`` (see its README); ``pow_mod``/``mod_inv`` there are CPython ``pow``.
* Ciphertexts are produced by ``oracle.paillier_oracle.encrypt_raw`` (third-party
  ``Paillier.encrypt`` is not available; its value-level parity is unpinned, the reference tests'
  own criterion decrypt(encrypt(m)) == m is recorded for every vector).
* Synthetic (dealer-simulated) keys come from ``oracle.keys.dealer_keygen``.
"""

from __future__ import annotations

import base64
import glob
import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path[:0] = [REF_SRC, os.path.join(HERE, "ref_shims"), ROOT]

from tno.mpc.encryption_schemes.paillier import PaillierCiphertext, PaillierPublicKey  # noqa: E402
from tno.mpc.encryption_schemes.shamir import IntegerShares, ShamirSecretSharingIntegers  # noqa: E402
from tno.mpc.protocols.distributed_keygen import DistributedPaillier, PaillierSharedKey  # noqa: E402
from tno.mpc.protocols.distributed_keygen.utils import AdditiveVariable, Batched  # noqa: E402

from oracle import keys as okeys  # noqa: E402
from oracle.paillier_oracle import encrypt_raw  # noqa: E402

FIXTURE_GLOB = os.path.join(
    REF_SRC, "tno/mpc/protocols/distributed_keygen/test/test_data/*.obj"
)


class _Scheme:
    """Just enough of a Paillier scheme for ``partial_decrypt``'s key check
    (``paillier_shared_key.py:67``)."""

    def __init__(self, n: int) -> None:
        self.public_key = PaillierPublicKey(n, n + 1)


def ref_key(n, t, player_id, shares, degree, scaling, parties, kappa, theta) -> PaillierSharedKey:
    scheme = ShamirSecretSharingIntegers(kappa, n, parties, t)
    return PaillierSharedKey(
        n=n,
        t=t,
        player_id=player_id,
        share=IntegerShares(scheme, dict(shares), degree, scaling),
        theta=theta,
    )


def ref_key_from_oracle_key(k) -> PaillierSharedKey:
    return ref_key(
        k.n, k.t, k.player_id, k.share.shares, k.share.degree, k.share.scaling,
        k.share.number_of_parties, k.share.kappa, k.theta,
    )


def decrypt_vectors(ref_keys: dict[int, PaillierSharedKey], n: int, rng: random.Random, count: int):
    """Run the reference's partial_decrypt for every party and decrypt for party 1."""
    scheme = _Scheme(n)
    plain = [1, -1, 2, -2, 3, -3, 150000000, -150000000, 4242424242, -4242424242, 0, n - 1]
    vectors = []
    for idx in range(count):
        m = plain[idx] if idx < len(plain) else rng.randrange(n)
        r = rng.randrange(1, n)
        c = encrypt_raw(n, m, r)
        partials = {
            pid: int(key.partial_decrypt(PaillierCiphertext(c, scheme)))
            for pid, key in ref_keys.items()
        }
        combined = {pid: int(key.decrypt(dict(partials))) for pid, key in ref_keys.items()}
        assert len(set(combined.values())) == 1 and combined[1] == m % n
        vectors.append(
            {
                "m": hex(m % n),
                "r": hex(r),
                "c": hex(c),
                "partials": {str(pid): hex(v) for pid, v in partials.items()},
                "plaintext": hex(combined[1]),
            }
        )
    # error path (paillier_shared_key.py:119-123): tampered partial -> ValueError
    bad = dict(partials)
    bad[1] = (bad[1] + 1) % (n * n)
    try:
        ref_keys[1].decrypt(bad)
        raised = False
    except ValueError:
        raised = True
    assert raised
    vectors.append(
        {
            "c": hex(c),
            "partials": {str(pid): hex(v) for pid, v in bad.items()},
            "error": "ValueError",
        }
    )
    # missing share (paillier_shared_key.py:108-110): KeyError
    return vectors


def make_fixture_vectors() -> dict:
    files = sorted(glob.glob(FIXTURE_GLOB))
    assert len(files) == 24, files
    sets: dict[tuple[int, int], dict] = {}
    for path in files:
        m = re.search(r"threshold_(\d)_(\d)parties_(\d)\.obj$", path)
        t, parties = int(m.group(1)), int(m.group(2))
        blob = open(path, "rb").read()
        k = okeys.key_from_blob(blob)
        entry = sets.setdefault((t, parties), {"t": t, "parties": parties, "keys": [], "_ref": {}})
        entry["keys"].append(
            {
                "file": os.path.basename(path),
                "blob_b64": base64.b64encode(blob).decode(),
                "n": hex(k.n),
                "theta": hex(k.theta),
                "player_id": k.player_id,
                "degree": k.share.degree,
                "scaling": k.share.scaling,
                "kappa": k.share.kappa,
                "share": hex(k.share.shares[k.player_id]),
            }
        )
        entry["_ref"][k.player_id] = ref_key_from_oracle_key(k)
    rng = random.Random(20261017)
    out = []
    for (t, parties), entry in sorted(sets.items()):
        refs = entry.pop("_ref")
        n = refs[1].n
        entry["vectors"] = decrypt_vectors(refs, n, rng, 16)
        out.append(entry)
    return {"source": "reference PaillierSharedKey on test/test_data/*.obj", "sets": out}


def make_dealer_vectors() -> dict:
    specs = [
        # name, key_length, parties, t, exact, seed, n_vectors
        ("cfg1_k512_p3_t1", 512, 3, 1, False, 20261018, 8),
        ("cfg2_k2048_p3_t1_exact", 2048, 3, 1, True, 20261019, 3),
        ("cfg2_k2048_p3_t1_real", 2048, 3, 1, False, 20261020, 3),
        ("cfg3_k2048_p5_t2_exact", 2048, 5, 2, True, 20261021, 2),
        ("cfg4_k4096_p3_t1_exact", 4096, 3, 1, True, 20261022, 1),
        ("small_k128_p3_t1", 128, 3, 1, False, 20261023, 8),
    ]
    out = {}
    for name, kl, parties, t, exact, seed, nvec in specs:
        print("dealer key", name, flush=True)
        dk = okeys.dealer_keygen(kl, parties, t, seed=seed, exact=exact)
        refs = {pid: ref_key_from_oracle_key(k) for pid, k in dk.keys.items()}
        rng = random.Random(seed + 1)
        out[name] = {
            "key_length": kl,
            "exact": exact,
            "key": okeys.dealer_key_to_json(dk),
            "vectors": decrypt_vectors(refs, dk.n, rng, nvec),
        }
    return {"source": "oracle.keys.dealer_keygen keys, reference PaillierSharedKey arithmetic", "keys": out}


DIGEST_SPECS = [
    # name (same keys as dealer_vectors.json: same seeds), key_length, parties, t, exact, seed, vectors
    ("cfg1_k512_p3_t1", 512, 3, 1, False, 20261018, 160),
    ("cfg2_k2048_p3_t1_exact", 2048, 3, 1, True, 20261019, 96),
    ("cfg2_k2048_p3_t1_real", 2048, 3, 1, False, 20261020, 96),
    ("cfg3_k2048_p5_t2_exact", 2048, 5, 2, True, 20261021, 64),
    ("cfg4_k4096_p3_t1_exact", 4096, 3, 1, True, 20261022, 8),
]


from digests import digest_inputs, sha  # noqa: E402  (tests/golden/digests.py)


def make_dealer_digests() -> dict:
    """More vectors per key than dealer_vectors.json holds in full (more than one warp of the
    2048/4096-bit kernels): inputs are regenerated from the seed, the reference's outputs are
    recorded as SHA-256 digests (ciphertext, every party's partial decryption, plaintext)."""
    out = {}
    for name, kl, parties, t, exact, seed, nvec in DIGEST_SPECS:
        print("digest set", name, nvec, flush=True)
        dk = okeys.dealer_keygen(kl, parties, t, seed=seed, exact=exact)
        refs = {pid: ref_key_from_oracle_key(k) for pid, k in dk.keys.items()}
        scheme = _Scheme(dk.n)
        rows = []
        for m, r in digest_inputs(dk.n, seed, nvec):
            c = encrypt_raw(dk.n, m, r)
            partials = {pid: int(key.partial_decrypt(PaillierCiphertext(c, scheme))) for pid, key in refs.items()}
            plain = int(refs[1].decrypt(dict(partials)))
            assert plain == m
            rows.append({"c": sha(c), "partials": {str(pid): sha(v) for pid, v in partials.items()}, "plaintext": sha(plain)})
        out[name] = {"key_length": kl, "parties": parties, "t": t, "exact": exact, "seed": seed, "n": hex(dk.n), "vectors": rows}
    return {"source": "oracle.keys.dealer_keygen keys, reference PaillierSharedKey arithmetic, SHA-256 of the values",
            "keys": out}


def make_biprime_vectors() -> dict:
    """Run the reference's v calculation and verdict on synthetic candidates: real biprimes (must
    pass) and random products of non-primes (must fail)."""
    v_calc = DistributedPaillier._DistributedPaillier__biprime_test_v_calculation
    verdict = DistributedPaillier._DistributedPaillier__biprime_test_with_v_i
    rng = random.Random(20261024)
    cases = []
    for key_length, parties, correct, want_biprime in [
        (64, 3, 20, True), (64, 3, 20, False), (64, 4, 20, False), (128, 3, 40, True),
        (128, 5, 40, False), (256, 3, 40, True), (256, 3, 40, False), (512, 3, 40, False),
        (2048, 3, 40, False),
    ]:
        pl = key_length // 2
        while True:
            p_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(parties)]
            q_sh = [okeys.prime_candidate_share(i + 1, pl, rng) for i in range(parties)]
            p, q = sum(p_sh), sum(q_sh)
            is_bp = okeys._is_probable_prime(p, rng) and okeys._is_probable_prime(q, rng)
            if is_bp == want_biprime:
                break
        n = p * q
        party_indices = {f"party{i}": i for i in range(1, parties + 1)}
        # jointly random g's: 4x as many as needed (distributed_keygen.py:1028)
        g_values = [rng.randint(0, n) % n for _ in range(correct * 4)]
        per_party = {}
        batched = {}
        for i in range(1, parties + 1):
            b = v_calc(g_values, i, n, p_sh[i - 1], q_sh[i - 1], correct)
            batched[i] = b
            per_party[i] = [int(v.get_share(i)) for v in b.variables if i in v._sharing]
        # assemble what exchange_reconstruct would leave behind: every party's share in party 1's
        # batched variable
        merged = Batched(AdditiveVariable(label="v", modulus=n), batch_size=correct)
        for i in range(1, parties + 1):
            merged.set_share(i, per_party[i])
        try:
            ok = bool(verdict(merged, n, correct, party_indices))
        except KeyError:
            ok = False  # fewer than `correct` usable g's: the reference would hit a missing share
        cases.append(
            {
                "key_length": key_length,
                "parties": parties,
                "correct_param_biprime": correct,
                "n": hex(n),
                "p_shares": [hex(x) for x in p_sh],
                "q_shares": [hex(x) for x in q_sh],
                "g_values": [hex(g) for g in g_values],
                "v": {str(i): [hex(v) for v in vs] for i, vs in per_party.items()},
                "verdict": ok,
                "is_biprime": is_bp,
            }
        )
        assert ok == is_bp or not is_bp, (key_length, parties, ok, is_bp)
    return {"source": "reference __biprime_test_v_calculation / __biprime_test_with_v_i", "cases": cases}


def main() -> None:
    for name, fn in [
        ("fixture_vectors.json", make_fixture_vectors),
        ("biprime_vectors.json", make_biprime_vectors),
        ("dealer_vectors.json", make_dealer_vectors),
        ("dealer_digests.json", make_dealer_digests),
    ]:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        data = fn()
        path = os.path.join(HERE, name)
        with open(path, "w") as fh:
            json.dump(data, fh, indent=0, separators=(",", ":"))
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
