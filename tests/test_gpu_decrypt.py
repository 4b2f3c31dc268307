"""GPU parity for share combination and the full threshold decryption, through the host mirror of
PaillierSharedKey, against the values recorded from the reference (tests/golden/*.json)."""
from __future__ import annotations

import base64
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _h(x: str) -> int:
    return int(x, 16)


def _gpu_keys(okeys_by_pid):
    from protocols.distributed_keygen_b200 import IntegerShares, PaillierSharedKey

    out = {}
    for pid, k in okeys_by_pid.items():
        share = IntegerShares(dict(k.share.shares), k.share.degree, k.share.scaling, k.share.number_of_parties)
        out[pid] = PaillierSharedKey(k.n, k.t, k.player_id, share, k.theta)
    return out


def _check_set(okeys_by_pid, vectors):
    keys = _gpu_keys(okeys_by_pid)
    good = [v for v in vectors if "error" not in v]
    cs = [_h(v["c"]) for v in good]
    partials = {pid: key.partial_decrypt_batch(cs) for pid, key in keys.items()}
    for pid in keys:
        assert partials[pid] == [_h(v["partials"][str(pid)]) for v in good]
    dicts = [{pid: partials[pid][i] for pid in keys} for i in range(len(good))]
    for pid, key in keys.items():
        assert key.decrypt_batch(dicts) == [_h(v["plaintext"]) for v in good]
        assert key.partial_decrypt_exponent() == okeys_by_pid[pid].partial_decrypt_exponent()
    # scalar API + error paths of the reference
    k1 = keys[1]
    assert k1.decrypt(dicts[0]) == _h(good[0]["plaintext"])
    assert k1.partial_decrypt(cs[0]) == partials[1][0]
    for v in vectors:
        if "error" in v:
            with pytest.raises(ValueError):
                k1.decrypt({int(p): _h(x) for p, x in v["partials"].items()})
    missing = dict(dicts[0])
    del missing[1]
    with pytest.raises(KeyError):
        k1.decrypt(missing)
    with pytest.raises(TypeError):
        k1.partial_decrypt("not a ciphertext")
    # status flags: a batch with one tampered element flags only that element
    ctx = k1._combine_ctx()
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints

    shares = k1.share.degree + 1
    arr = np.stack([ints_to_limbs([d[i + 1] for d in dicts], ctx.n2_limbs) for i in range(shares)])
    if arr.shape[1] >= 2:
        arr[0, 1, 0] ^= 1
        out, status = ctx.combine_limbs(arr)
        assert status[1] == 2 and not status[[i for i in range(arr.shape[1]) if i != 1]].any()
        assert limbs_to_ints(out)[0] == _h(good[0]["plaintext"]) and limbs_to_ints(out)[1] == 0
    for key in keys.values():
        key.close()


def test_threshold_decrypt_fixture_keys(fixture_vectors):
    from oracle import keys as okeys

    for entry in fixture_vectors["sets"]:
        ok = {k["player_id"]: okeys.key_from_blob(base64.b64decode(k["blob_b64"])) for k in entry["keys"]}
        _check_set(ok, entry["vectors"])


def test_threshold_decrypt_dealer_keys(dealer_vectors):
    from oracle import keys as okeys

    for name, item in dealer_vectors["keys"].items():
        dk = okeys.dealer_key_from_json(item["key"])
        _check_set(dk.keys, item["vectors"])


def test_decrypt_sequence_roundtrip_cfg1():
    """BASELINE config 1 shape: 3 parties, t=1, key_length 512, a sequence of ciphertexts:
    decrypt(encrypt(m)) == m (the reference tests' criterion, test_distributed_keygen.py:161-185)
    and bit-exact partials against the oracle."""
    from oracle import keys as okeys
    from oracle.paillier_oracle import encrypt_raw

    dk = okeys.dealer_keygen(512, 3, 1, seed=20261018)
    keys = _gpu_keys(dk.keys)
    rng = random.Random(3)
    ms = [rng.randrange(-(2**40), 2**40) for _ in range(1000)]
    cs = [encrypt_raw(dk.n, m, rng.randrange(1, dk.n)) for m in ms]
    partials = {pid: key.partial_decrypt_batch(cs) for pid, key in keys.items()}
    for pid in (1, 2, 3):
        sample = rng.sample(range(1000), 25)
        for i in sample:
            assert partials[pid][i] == dk.keys[pid].partial_decrypt(cs[i])
    dicts = [{pid: partials[pid][i] for pid in keys} for i in range(len(cs))]
    assert keys[2].decrypt_batch(dicts) == [m % dk.n for m in ms]
    for key in keys.values():
        key.close()


def test_decrypt_through_wire_messages(dealer_vectors):
    """Section 8 f4: ciphertext limb rows -> partial decryption -> message bytes -> combination,
    no Python int on the way; bytes equal a generic msgpack serialisation of the oracle's values
    with the reference's big-integer tagging, plaintexts equal the reference's."""
    import msgpack
    from oracle import keys as okeys
    from protocols.distributed_keygen_b200 import distributed_keygen as dkg
    from protocols.distributed_keygen_b200.limbs import ints_to_limbs, limbs_to_ints
    from protocols.distributed_keygen_b200.paillier_shared_key import n_square_limbs

    item = dealer_vectors["keys"]["cfg2_k2048_p3_t1_real"]
    dk = okeys.dealer_key_from_json(item["key"])
    keys = _gpu_keys(dk.keys)
    good = [v for v in item["vectors"] if "error" not in v]
    cs = [_h(v["c"]) for v in good]
    rows = ints_to_limbs(cs, n_square_limbs(dk.n))
    messages = {pid: dkg.partial_decryption_message(key, rows) for pid, key in keys.items()}
    for pid in keys:
        want = [_h(v["partials"][str(pid)]) for v in good]
        generic = msgpack.packb(
            {"content": "partial_decryption_sequence",
             "value": [{"type": "int", "data": w.to_bytes((w.bit_length() + 8) // 8, "little", signed=True)} for w in want]},
            use_bin_type=True)
        assert messages[pid] == generic
    out = dkg.decrypt_from_messages(keys[3], messages)
    assert limbs_to_ints(out) == [_h(v["plaintext"]) for v in good]
    with pytest.raises(KeyError):
        dkg.decrypt_from_messages(keys[1], {1: messages[1], 3: messages[3]})
    tampered = dict(messages)
    other = ints_to_limbs([c + 1 for c in cs], n_square_limbs(dk.n))
    tampered[2] = dkg.partial_decryption_message(keys[2], other)
    with pytest.raises(ValueError, match="not divisible by N"):
        dkg.decrypt_from_messages(keys[1], tampered)
    for key in keys.values():
        key.close()
