"""The oracle against the golden vectors recorded from the reference's own classes
(tests/golden/make_golden.py), against GMP, and against the reference test-suite's criterion
decrypt(encrypt(m)) == m (src/test/test_distributed_keygen.py:111-129)."""
from __future__ import annotations

import base64
import random

import pytest

from oracle import gmp
from oracle import keys as okeys
from oracle import paillier_oracle as po


def _h(x: str) -> int:
    return int(x, 16)


def _fixture_keys(entry):
    return {k["player_id"]: okeys.key_from_blob(base64.b64decode(k["blob_b64"])) for k in entry["keys"]}


def test_fixture_blobs_decode(fixture_vectors):
    assert len(fixture_vectors["sets"]) == 6
    total = 0
    for entry in fixture_vectors["sets"]:
        for k in entry["keys"]:
            key = okeys.key_from_blob(base64.b64decode(k["blob_b64"]))
            assert key.n == _h(k["n"]) and key.theta == _h(k["theta"])
            assert key.player_id == k["player_id"] and key.share.degree == k["degree"] == 2 * entry["t"]
            assert key.share.shares[key.player_id] == _h(k["share"])
            assert key.share.scaling == k["scaling"]
            total += 1
    assert total == 24


def test_oracle_matches_reference_on_fixture_keys(fixture_vectors):
    """partial_decrypt / decrypt values recorded from the reference's PaillierSharedKey."""
    for entry in fixture_vectors["sets"]:
        keys = _fixture_keys(entry)
        for vec in entry["vectors"]:
            c = _h(vec["c"])
            partials = {int(p): _h(v) for p, v in vec["partials"].items()}
            if "error" in vec:
                with pytest.raises(ValueError):
                    keys[1].decrypt(partials)
                continue
            for pid, key in keys.items():
                assert key.partial_decrypt(c) == partials[pid]
                assert key.decrypt(partials) == _h(vec["plaintext"])
            n = keys[1].n
            assert po.encrypt_raw(n, _h(vec["m"]), _h(vec["r"])) == c
            assert _h(vec["plaintext"]) == _h(vec["m"])


def test_oracle_matches_reference_on_dealer_keys(dealer_vectors):
    for name, item in dealer_vectors["keys"].items():
        dk = okeys.dealer_key_from_json(item["key"])
        assert dk.p * dk.q == dk.n
        for vec in item["vectors"]:
            partials = {int(p): _h(v) for p, v in vec["partials"].items()}
            if "error" in vec:
                with pytest.raises(ValueError):
                    dk.keys[1].decrypt(partials)
                continue
            c = _h(vec["c"])
            # GMP ("gmpy2 path") and CPython pow must agree with the recorded reference values
            for pid, key in dk.keys.items():
                assert key.partial_decrypt(c) == partials[pid]
                assert gmp.powm(c, key.partial_decrypt_exponent(), key.n_square) == partials[pid]
            assert dk.keys[1].decrypt(partials) == _h(vec["plaintext"]) == _h(vec["m"])


def test_missing_share_is_keyerror(fixture_vectors):
    """paillier_shared_key.py:108-110 indexes the dict: a missing party is a KeyError."""
    entry = fixture_vectors["sets"][3]
    keys = _fixture_keys(entry)
    partials = {int(p): _h(v) for p, v in entry["vectors"][0]["partials"].items()}
    del partials[1]
    with pytest.raises(KeyError):
        keys[2].decrypt(partials)


def test_biprime_v_matches_reference(biprime_vectors):
    for case in biprime_vectors["cases"]:
        n = _h(case["n"])
        g_values = [_h(g) for g in case["g_values"]]
        correct = case["correct_param_biprime"]
        v_by_party = {}
        for i in range(1, case["parties"] + 1):
            p_i, q_i = _h(case["p_shares"][i - 1]), _h(case["q_shares"][i - 1])
            v = po.biprime_v_calculation(g_values, i, n, p_i, q_i, correct)
            assert v == [_h(x) for x in case["v"][str(i)]]
            v_by_party[i] = v
        if all(len(v) >= correct for v in v_by_party.values()):
            assert po.biprime_verdict(v_by_party, n, correct) == case["verdict"]
        assert case["verdict"] == case["is_biprime"]


def test_jacobi_against_gmp():
    rng = random.Random(7)
    for _ in range(300):
        n = rng.getrandbits(rng.choice([16, 64, 300])) | 1
        a = rng.getrandbits(310)
        assert po.jacobi(a, n) == gmp.jacobi(a, n)


def test_gmp_and_cpython_agree_with_negative_exponents():
    rng = random.Random(11)
    dk = okeys.dealer_keygen(128, 3, 1, seed=99)
    n2 = dk.n * dk.n
    for _ in range(20):
        c = rng.randrange(1, n2)
        e = rng.getrandbits(300) * rng.choice([1, -1])
        assert gmp.powm(c, e, n2) == pow(c, e, n2)
    with pytest.raises(ZeroDivisionError):
        po.mod_inv(dk.p, n2)
    with pytest.raises(ZeroDivisionError):
        gmp.invert(dk.p, n2)


def test_reference_shaped_key_roundtrip_cfg1():
    """BASELINE config 1: 3 parties, t=1, key_length 512, decrypt on CPU (sampled)."""
    dk = okeys.dealer_keygen(512, 3, 1, seed=20261018)
    assert 513 <= dk.n.bit_length() <= 516
    rng = random.Random(5)
    for _ in range(25):
        m = rng.randrange(-(2**32), 2**32)
        c = po.encrypt_raw(dk.n, m, rng.randrange(1, dk.n))
        partials = {i: k.partial_decrypt(c) for i, k in dk.keys.items()}
        assert dk.keys[2].decrypt(partials) == m % dk.n


def test_mult_list():
    assert po.mult_list([]) == 1
    assert po.mult_list([3, 5, 7]) == 105
    assert po.mult_list([3, 5, 7], 11) == 105 % 11
